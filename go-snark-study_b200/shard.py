"""Work sharding of a Groth16 proving key across ranks — the same arithmetic as
csrc/shard_partition.h, which groth16_pk_load uses; tests/test_shard_partition.py
holds the two against each other.

The four MSMs (A, B1, B2, C||PTD) are laid end to end on a line weighted by cost (a G2
term ~ 2.8 G1 terms).  With no phase cost, rank g takes the g-th of `world` equal pieces.
With a phase cost, every piece a rank holds is one more MSM on that rank (its own digit
sort, slice merge and bucket tail), so opening a piece of set k costs FIX[k] on top of the
per-term weight: the ranks are filled greedily up to a common capacity T, the smallest one
that covers the line.  A rank therefore holds whole MSMs where it can and index ranges
where it must.  A set's blinding points (alpha/beta/delta tails) belong to the rank that
holds the set's last element.
"""

W_AB, W_G2 = 1.0, 2.8          # mirror of g_w_ab / g_w_g2 (csrc/capi.cu)
PHASE_COST = 0.0               # mirror of g_phase_cost (B200_CFG_SHARD_PHASE_COST): G1 terms per piece; a G2 piece counts twice


def weights(world):
    """A, B1, B2 (G2), C||PTD in G1 terms of the C||PTD set (csrc/prove_host.cuh: groth16_pk_load)."""
    w_ab = W_AB if world > 1 else 1.0
    return (w_ab, w_ab, W_G2, 1.0)


def phase_costs(world):
    f = PHASE_COST if world > 1 else 0.0
    return (f, f, 2.0 * f, f)


def _sweep(lens, wgt, fix, world, T, force_last):
    """csrc/shard_partition.h: shard_sweep."""
    k, pos, cuts = 0, 0, []
    for g in range(world):
        lo = [lens[j] if j < k else (pos if j == k else 0) for j in range(4)]
        hi = list(lo)
        cap = T
        unbounded = force_last and g == world - 1
        while k < 4:
            rem = lens[k] - pos
            if rem == 0:
                k, pos = k + 1, 0
                continue
            take = rem
            if not unbounded:
                if cap <= fix[k]:
                    break
                avail = (cap - fix[k]) / wgt[k]
                if avail < float(rem):
                    take = int(avail)
                if take == 0:
                    break
            lo[k], hi[k] = pos, pos + take
            cap -= fix[k] + wgt[k] * float(take)
            pos += take
            if pos < lens[k]:
                break
            k, pos = k + 1, 0
        cuts.append((lo, hi))
    return k == 4, cuts


def partition(lens, wgt, fix, world):
    """csrc/shard_partition.h: shard_partition -> [(lo[4], hi[4])] per rank."""
    off = [0.0]
    for w, ln in zip(wgt, lens):
        off.append(off[-1] + w * float(ln))
    if all(f == 0 for f in fix):
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
    t_lo, t_hi = 0.0, off[4] + fix[0] + fix[1] + fix[2] + fix[3] + 1.0
    for _ in range(64):
        mid = 0.5 * (t_lo + t_hi)
        if _sweep(lens, wgt, fix, world, mid, False)[0]:
            t_hi = mid
        else:
            t_lo = mid
    return _sweep(lens, wgt, fix, world, t_hi, True)[1]


def shard_ranges(m, npublic, n_ptd, rank, world):
    assert world >= 1 and 0 <= rank < world
    n_c_full = m - npublic - 1
    lens = (m, m, m, n_c_full + n_ptd)
    lo, hi = partition(lens, weights(world), phase_costs(world), world)[rank]
    out = {"n_c_full": n_c_full, "sets": []}
    for k in range(4):
        tail = hi[k] == lens[k] and (lo[k] < hi[k] or (rank == world - 1 and lens[k] == 0))
        out["sets"].append({"lo": lo[k], "hi": hi[k], "tail": tail})
    return out
