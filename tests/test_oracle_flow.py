"""CPU tests (no GPU) for the consumers of the vector clips (mv.SCDetection / mv.Mask / mv.Finest /
mv.Flow): the oracle restatement (oracle/mvtools_oracle_flow.c) is pinned against the reference.

 * test_flow_oracle_matches_golden: against committed hashes of the reference's own outputs
   (tests/golden/golden_flow.json, made by tests/golden/make_golden_flow.py from oracle/_ref);
 * test_flow_oracle_matches_compiled_reference: live against oracle/_ref/libmvtools_ref.so where present;
 * create-time logic: error texts and derived values of mvb200_mask_init / _flow_init / _scdetection_init,
   and the SimpleResize tables against the oracle's."""
import json
import os

import numpy as np
import pytest

import mvtools_b200 as mb
import pipeline
import pipeline_flow

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_flow.json")))


def oracle_run(case):
    import oracle
    out = pipeline_flow.run_engine_flow(case, oracle.scdetection_frame, oracle.mask_frame, oracle.finest_frame, oracle.flow_frame)
    out["inter"] = pipeline_flow.run_engine_inter(case, oracle.flowinter_frame, oracle.blend_frame)
    out["fps"] = pipeline_flow.run_engine_fps(case, oracle.flowinter_frame, oracle.blend_frame)
    out["blur"] = pipeline_flow.run_engine_blur(case, oracle.flowblur_frame)
    out["bfps"] = pipeline_flow.run_engine_bfps(case, oracle.blockfps_frame, oracle.blend_frame)
    return out


@pytest.mark.parametrize("case", sorted(pipeline_flow.FLOW))
def test_flow_oracle_matches_golden(built, case):
    got = oracle_run(case)
    for stage in ("scd", "mask", "flow", "finest", "inter", "fps", "blur", "bfps"):
        assert got[stage] == GOLD[case][stage], "%s: oracle %s differs from the reference's golden hashes" % (case, stage)


@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "p10_444_blk8_pel4_bicubic", "noise_badblocks_blk8"])
def test_flow_oracle_matches_compiled_reference(have_ref, case):
    if not have_ref:
        pytest.skip("oracle/_ref/libmvtools_ref.so not built here (no /root/reference)")
    ref = pipeline_flow.run_reference_flow(case)
    assert ref == GOLD[case], "committed goldens are stale: re-run tests/golden/make_golden_flow.py"
    assert oracle_run(case) == ref


def test_flow_goldens_exercise_both_paths():
    """the pinned cases contain usable and unusable frames, and scene-change flags of both values"""
    flags = [f for case in GOLD.values() for variant in case["scd"] for f in variant]
    assert 0 in flags and 1 in flags


def test_simple_resize_tables(built):
    import oracle
    for out, in_ in ((320, 79), (192, 47), (640, 80), (17, 2), (100, 100), (7, 3), (1920, 239), (1080, 134)):
        a = mb.simple_resize_tables(out, in_)
        b = oracle.simple_resize_tables(out, in_)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (out, in_)
        assert a[0].min() >= 0 and a[0].max() <= in_ - 2 and a[1].min() >= 0 and a[1].max() <= 1 << 14


def _vd(case="blk8_ov4_pel2_p8"):
    sp, aps, dp = pipeline.params_for(case)
    return sp, aps[0].ad


def test_mask_init_errors_and_values(built):
    sp, vd = _vd()
    for kw, msg in ((dict(gamma=-1.0), "Mask: gamma must not be negative."),
                    (dict(kind=6), "Mask: kind must 0, 1, 2, 3, 4, or 5."),
                    (dict(time=100.5), "Mask: time must be between 0.0 and 100.0 (inclusive)."),
                    (dict(ysc=256), "Mask: ysc must be between 0 and 255 (inclusive)."),
                    (dict(thscd1=16321), "Mask: thscd1 can be at most 16320.")):
        with pytest.raises(mb.Error) as e:
            mb.mask_params(vd, **kw)
        assert str(e.value) == msg
    with pytest.raises(mb.Error) as e:
        mb.mask_params(vd, clip_bits=10)
    assert str(e.value) == "Mask: input clip must be GRAY8, YUV420P8, YUV422P8, YUV440P8, or YUV444P8, with constant dimensions."
    mp = mb.mask_params(vd, time=33.0, ml=50.0)
    assert mp.time256 == int(33.0 * 256 / 100) and mp.fMaskNormFactor == np.float32(1.0) / np.float32(50.0)
    assert (mp.nWidthB, mp.nHeightB) == (vd.nBlkX * (vd.nBlkSizeX - vd.nOverlapX) + vd.nOverlapX, vd.nBlkY * (vd.nBlkSizeY - vd.nOverlapY) + vd.nOverlapY)


def test_flow_init_errors_and_values(built):
    sp, vd = _vd()
    for kw, msg in ((dict(time=-1.0), "Flow: time must be between 0 and 100 % (inclusive)."),
                    (dict(mode=2), "Flow: mode must be 0 or 1."),
                    (dict(thscd1=20000), "Flow: thscd1 can be at most 16320.")):
        with pytest.raises(mb.Error) as e:
            mb.flow_params(sp, vd, 6, **kw)
        assert str(e.value) == msg
    sp2, _ = _vd("c1_640x360_p8_blk8_pel1")
    with pytest.raises(mb.Error) as e:
        mb.flow_params(sp2, vd, 6)
    assert str(e.value) == "Flow: wrong source or super clip frame size."
    fp = mb.flow_params(sp, vd, 6, time=50.0)
    assert fp.time256 == 128 and fp.nBlkXP >= vd.nBlkX and fp.nWidthP >= vd.nWidth and fp.VPitchY % 16 == 0


def test_scdetection_init(built):
    sp, vd = _vd()
    p = mb.scdetection_params(vd)
    assert p.nSCD2 == 130 * vd.nBlkX * vd.nBlkY // 256
    with pytest.raises(mb.Error) as e:
        mb.scdetection_params(vd, thscd1=99999)
    assert str(e.value) == "SCDetection: thscd1 can be at most 16320."


def _two_vds(case="blk8_ov4_pel2_p8"):
    sp, aps, dp = pipeline.params_for(case)
    return sp, aps[0].ad, aps[1].ad


def test_interpolation_filters_init_errors(built):
    """create-time checks of mv.FlowInter / mv.FlowFPS / mv.FlowBlur / mv.BlockFPS: the reference's texts, in its order"""
    sp, bw, fw = _two_vds()
    W, H, N = 320, 192, 6
    cases = [
        (lambda: mb.flowinter_params(sp, bw, fw, N, time=100.5), "FlowInter: time must be between 0 and 100 % (inclusive)."),
        (lambda: mb.flowinter_params(sp, bw, fw, N, ml=0.0), "FlowInter: ml must be greater than 0."),
        (lambda: mb.flowinter_params(sp, fw, fw, N), "FlowInter: mvbw must be generated with isb=True."),
        (lambda: mb.flowinter_params(sp, bw, bw, N), "FlowInter: mvfw must be generated with isb=False."),
        (lambda: mb.flowinter_params(sp, bw, fw, N, thscd1=17000), "FlowInter: thscd1 can be at most 16320."),
        (lambda: mb.flowfps_params(sp, bw, fw, N, W, H, mask=3), "FlowFPS: mask must be 0, 1, or 2."),
        (lambda: mb.flowfps_params(sp, bw, fw, N, W, H, ml=-1.0), "FlowFPS: ml must be greater than 0."),
        (lambda: mb.flowfps_params(sp, bw, fw, N, W, H, fps_num=0), "FlowFPS: The input clip must have a frame rate. Invoke AssumeFPS if necessary."),
        (lambda: mb.flowfps_params(sp, bw, fw, N, W + 2, H), "FlowFPS: inconsistent source and vector frame size."),
        (lambda: mb.flowfps_params(sp, fw, bw, N, W, H), "FlowFPS: mvbw must be generated with isb=True."),
        (lambda: mb.flowblur_params(sp, bw, fw, N, blur=200.5), "FlowBlur: blur must be between 0 and 200 % (inclusive)."),
        (lambda: mb.flowblur_params(sp, bw, fw, N, prec=0), "FlowBlur: prec must be at least 1."),
        (lambda: mb.flowblur_params(sp, bw, bw, N), "FlowBlur: mvfw must be generated with isb=False."),
        (lambda: mb.blockfps_params(sp, bw, fw, N, W, H, mode=9), "BlockFPS: mode must be between 0 and 8 (inclusive)."),
        (lambda: mb.blockfps_params(sp, bw, fw, N, W, H + 2), "BlockFPS: wrong source or super clip frame size."),
        (lambda: mb.blockfps_params(sp, bw, fw, N, W, H, fps_den=0), "BlockFPS: The input clip must have a frame rate. Invoke AssumeFPS if necessary."),
    ]
    for call, msg in cases:
        with pytest.raises(mb.Error) as e:
            call()
        assert str(e.value) == msg
    sp16, bw16, fw16 = _two_vds("blk16_ov8_pel2_p16_r3")
    with pytest.raises(mb.Error) as e:
        mb.flowinter_params(sp, bw, fw16, N)
    assert str(e.value) == "FlowInter: mvbw and mvfw have different block sizes." or str(e.value).startswith("FlowInter: mvbw and mvfw have different")


def test_fps_conversion_arithmetic(built):
    """fa / fb, the output length and rate, and the position of every output frame (MVFlowFPS.c:118-141,737-756)"""
    sp, bw, fw = _two_vds()
    for (num, den, fn, fd, n_in) in ((50, 1, 25, 1, 6), (60, 1, 25, 1, 6), (30000, 1001, 25, 1, 9), (24, 1, 30000, 1001, 17), (0, 0, 25, 1, 5)):
        p = mb.flowfps_params(sp, bw, fw, n_in, 320, 192, fps_num=fn, fps_den=fd, num=num, den=den)
        numerator, denominator = (num, den) if num and den else (fn * 2, fd)
        from math import gcd
        fa, fb = denominator * fn, numerator * fd
        g = gcd(fa, fb)
        fa, fb = fa // g, fb // g
        assert (p.fa, p.fb) == (fa, fb)
        assert p.num_frames_out == 1 + (n_in - 1) * fb // fa
        g2 = gcd(numerator, denominator)
        assert (p.fps_num, p.fps_den) == (numerator // g2, denominator // g2)
        q = mb.blockfps_params(sp, bw, fw, n_in, 320, 192, fps_num=fn, fps_den=fd, num=num, den=den)
        assert (q.fa, q.fb, q.num_frames_out) == (p.fa, p.fb, p.num_frames_out)
        for n in range(p.num_frames_out):
            nleft, t = mb.flowfps_position(p, n)
            assert nleft == n * fa // fb and t == int((n * fa / fb - nleft) * 256 + 0.5)
            assert mb.blockfps_position(q, n) == (nleft, t)
