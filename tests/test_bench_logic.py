"""CPU tests of bench.py's host-side logic: the script builder, the cyclic clip, the shard plan and the output check against
tests/golden/bench_digest.json -- driven through the product's plugin glue over the stand-in engine (tests/mock), which
forwards to the oracle: the digests the reference produced at full size must come out for every rank's shard."""
import os

import pytest

import bench
import sharding
from test_plugin_glue import mock_plugin  # noqa: F401  (fixture)


def test_configs_name_the_baseline_configurations():
    assert set(bench.CONFIGS) == {"headline", "c1", "c2", "c3", "c4", "c5"}
    for name, cfg in bench.CONFIGS.items():
        assert cfg["kind"] in ("degrain", "analyse", "mc") and cfg["frames"] > 0 and cfg["metric"].startswith("frames/sec")
    assert bench.CONFIGS["headline"]["metric"] == bench.METRIC
    assert bench.CONFIGS["c5"]["fixed_total"] == 512  # the strong-scaling batch of BASELINE.json's configs[4]


def test_cyclic_clip_is_continuous_across_shards():
    cfg = bench.CONFIGS["c1"]
    whole = bench.make_frames(cfg, 70)
    part = bench.make_frames(cfg, 20, first=41)
    assert all(a is b for a, b in zip(whole[41:61], part))
    assert whole[0] is whole[30] and whole[29] is whole[59]


@pytest.mark.parametrize("world", [1, 2])
def test_output_check_matches_reference_digests_for_every_shard(mock_plugin, world, monkeypatch):  # noqa: F811
    """c1 (640x360, Super + Analyse) through the plugin glue: every rank's sampled frames equal the reference-derived digests
    (head frames on rank 0, cycle frames elsewhere), and a corrupted expectation is reported"""
    import minivs
    monkeypatch.setenv("MVB200_LOOKAHEAD", "8")
    gold = bench.load_golden("c1")
    assert gold is not None and len(gold["cycle"]) == 30 and len(gold["head"]) == bench.REACH
    cfg = bench.CONFIGS["c1"]
    F = 40
    for rank in range(world):
        plan = sharding.local_plan(world * F, world, rank, reach=bench.REACH)
        core = minivs.Core()
        mv = core.load(mock_plugin)
        out, keep = bench.build_script(mv, core, cfg, bench.make_frames(cfg, plan["local_frames"], first=plan["lo"]))
        seen = 0
        for n in range(plan["offset"], plan["offset"] + plan["count"]):
            want = bench.expected_digest(gold, plan["lo"] + n, world * F)
            if want is None:
                assert plan["lo"] + n >= world * F - bench.REACH
                continue
            if n % 5 == 0 or plan["lo"] + n < 4:
                assert bench.frame_digest(cfg, out.get_frame(n)) == want, "rank %d local frame %d (global %d)" % (rank, n, plan["lo"] + n)
                seen += 1
        assert seen >= 6
    assert bench.expected_digest(gold, 0, 100) == gold["head"][0] and bench.expected_digest(gold, 33, 100) == gold["cycle"][3]
    assert bench.expected_digest(gold, 98, 100) is None


def test_golden_digest_file_covers_the_default_workload():
    gold = bench.load_golden("headline")
    assert gold is not None and len(gold["cycle"]) == bench.DISTINCT and gold["workload"] == bench.CONFIGS["headline"]["workload"]
    assert os.path.exists(os.path.join(bench.ROOT, "tests", "golden", "make_golden_bench.py"))
