"""CPU: the index-range partition of a sharded Groth16 key (csrc/shard_partition.h, compiled with g++ behind a C entry
point) against its Python mirror (gosnark_b200/shard.py), and the properties the multi-GPU decomposition rests on:
every set is tiled exactly once in rank order, no rank's modelled load exceeds the common capacity the bisection found,
and with a phase cost no rank opens a piece it cannot pay for."""
import ctypes
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def cpart():
    import build
    lib = ctypes.CDLL(build.build_shard_partition())
    lib.shard_partition_c.restype = ctypes.c_int

    def run(lens, wgt, fix, world):
        l = np.array(lens, dtype=np.uint64)
        w = np.array(wgt, dtype=np.float64)
        f = np.array(fix, dtype=np.float64)
        out = np.zeros(8 * world, dtype=np.uint64)
        lib.shard_partition_c(l.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p),
                              f.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(world), out.ctypes.data_as(ctypes.c_void_p))
        o = out.reshape(world, 4, 2)
        return [([int(o[g, k, 0]) for k in range(4)], [int(o[g, k, 1]) for k in range(4)]) for g in range(world)]
    return run


def cases():
    rng = random.Random(11)
    out = [((1 << 20) + 2, 1, (1 << 20) + 1, w, f) for w in (1, 2, 3, 4, 8) for f in (0.0, 80000.0, 160000.0, 240000.0)]
    out += [((1 << 16) + 2, 1, (1 << 16) + 1, w, f) for w in (2, 8) for f in (0.0, 5000.0, 160000.0)]
    out += [(66, 1, 65, w, f) for w in (2, 3, 8) for f in (0.0, 3.0, 160000.0)]
    for _ in range(60):
        m = rng.randrange(3, 5000)
        out.append((m, rng.randrange(0, min(m - 1, 4)), rng.randrange(1, 5000), rng.randrange(1, 17), rng.choice((0.0, 1.0, 37.5, 900.0, 1e6))))
    return out


@pytest.mark.parametrize("w_ab,w_g2", [(1.0, 2.8), (1.3, 3.3)])
def test_cpp_partition_equals_python_mirror_and_tiles(cpart, w_ab, w_g2):
    from gosnark_b200 import shard
    for m, npublic, n_ptd, world, f in cases():
        lens = (m, m, m, m - npublic - 1 + n_ptd)
        wgt = (w_ab, w_ab, w_g2, 1.0)
        fix = (f, f, 2.0 * f, f)
        got = cpart(lens, wgt, fix, world)
        exp = shard.partition(lens, wgt, fix, world)
        assert got == [(list(lo), list(hi)) for lo, hi in exp], (m, world, f)
        # exact tiling in rank order
        for k in range(4):
            pos = 0
            for lo, hi in got:
                assert lo[k] <= hi[k] <= lens[k]
                if lo[k] < hi[k]:
                    assert lo[k] == pos
                    pos = hi[k]
            assert pos == lens[k], (m, world, f, k)
        # modelled load: the largest is what the bisection minimised — no rank but (by rounding) the last exceeds it by more
        # than one term's weight, and the total is conserved
        load = [sum((fix[k] + wgt[k] * (hi[k] - lo[k])) for k in range(4) if hi[k] > lo[k]) for lo, hi in got]
        pieces = sum(1 for lo, hi in got for k in range(4) if hi[k] > lo[k])
        assert abs(sum(load) - (sum(w * l for w, l in zip(wgt, lens)) + sum(fix[k] for lo, hi in got for k in range(4) if hi[k] > lo[k]))) < 1e-6 * (1 + sum(load))
        assert pieces <= 4 + world - 1
        if f == 0.0:
            ideal = sum(w * l for w, l in zip(wgt, lens)) / world
            assert max(load) <= ideal + 4 * max(wgt) + 1e-9


def test_shard_ranges_defaults_are_the_round2_partition():
    """PHASE_COST = 0 (the library default) reproduces the equal-pieces cuts recorded in profiles/r2_scale_n8_final.json."""
    from gosnark_b200 import shard
    assert shard.PHASE_COST == 0.0
    m = (1 << 20) + 2
    s1 = shard.shard_ranges(m, 1, (1 << 20) + 1, 1, 8)["sets"]
    assert (s1[0]["lo"], s1[0]["hi"], s1[1]["lo"], s1[1]["hi"]) == (891290, 1048578, 0, 734003)
    s7 = shard.shard_ranges(m, 1, (1 << 20) + 1, 7, 8)["sets"]
    assert (s7[3]["lo"], s7[3]["hi"], s7[3]["tail"]) == (1205862, 2097153, True)
