"""Work sharding of a Groth16 proving key across ranks — the same arithmetic as
csrc/shard_partition.h, which groth16_pk_load uses; tests/test_shard_partition.py
holds the two against each other.

The four MSMs (A, B1, B2, C||PTD) are laid end to end on a line weighted by cost (a G2
term ~ 2.8 G1 terms); rank g takes the g-th of `world` equal pieces.  A rank therefore
holds whole MSMs where it can and index ranges where it must.  A set's blinding points
(alpha/beta/delta tails) belong to the rank that holds the set's last element.
"""

W_AB, W_G2 = 1.0, 2.8          # mirror of g_w_ab / g_w_g2 (csrc/capi.cu)


def weights(world):
    """A, B1, B2 (G2), C||PTD in G1 terms of the C||PTD set (csrc/prove_host.cuh: groth16_pk_load)."""
    w_ab = W_AB if world > 1 else 1.0
    return (w_ab, w_ab, W_G2, 1.0)


def partition(lens, wgt, world):
    """csrc/shard_partition.h: shard_partition -> [(lo[4], hi[4])] per rank."""
    off = [0.0]
    for w, ln in zip(wgt, lens):
        off.append(off[-1] + w * float(ln))

    def cut(g, k):
        if g >= world:
            return lens[k]
        pos = off[4] * float(g) / float(world)
        x = (pos - off[k]) / wgt[k]
        if x <= 0:
            return 0
        if x >= float(lens[k]):
            return lens[k]
        return int(x)
    return [([cut(g, k) for k in range(4)], [cut(g + 1, k) for k in range(4)]) for g in range(world)]


def shard_ranges(m, npublic, n_ptd, rank, world):
    assert world >= 1 and 0 <= rank < world
    n_c_full = m - npublic - 1
    lens = (m, m, m, n_c_full + n_ptd)
    lo, hi = partition(lens, weights(world), world)[rank]
    out = {"n_c_full": n_c_full, "sets": []}
    for k in range(4):
        tail = hi[k] == lens[k] and (lo[k] < hi[k] or (rank == world - 1 and lens[k] == 0))
        out["sets"].append({"lo": lo[k], "hi": hi[k], "tail": tail})
    return out
