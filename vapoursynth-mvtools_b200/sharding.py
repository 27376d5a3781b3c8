"""Frame sharding for multi-GPU runs.

Frames are independent units of the hot path (no temporal recursion, readme.rst:36): a clip of N
frames is split into contiguous ranges, one per rank.  DegrainN of frame n reads super frames
n +- delta for delta <= radius, so a rank also needs a HALO of `reach` = radius * max_delta source
frames on each side of its range (MVDegrains.cpp:98-106, MVAnalyse.c:84-98); it recomputes their super
frames locally (Super is ~4 % of the step) instead of fetching them from the neighbour.  No data-path
collective is needed; the only communication of a job is the final gather of the output planes."""


def shard_range(num_frames, world, rank):
    """Contiguous, balanced split: returns (first, count) of `rank`'s frames."""
    base, extra = divmod(num_frames, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def with_halo(first, count, num_frames, reach):
    """The frame range a rank must hold locally: (lo, hi) with lo <= first, hi >= first + count."""
    return max(0, first - reach), min(num_frames, first + count + reach)


def local_plan(num_frames, world, rank, reach):
    """Everything a rank needs: its output range, its local range and where the output range sits in it."""
    first, count = shard_range(num_frames, world, rank)
    lo, hi = with_halo(first, count, num_frames, reach)
    return {"first": first, "count": count, "lo": lo, "hi": hi, "offset": first - lo, "local_frames": hi - lo}
