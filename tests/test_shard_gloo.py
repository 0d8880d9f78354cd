"""CPU, world_size 2 and 3 over gloo: the N>1 host logic — index-range shards cover
every term exactly once, and gathering per-rank partial sums (here in the exponent:
integers mod r stand in for the 1 KB XYZZ records) then adding them reproduces the
unsharded Groth16 discrete logs, blinding terms counted once (rank 0)."""
import os
import random

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _worker(rank, world, port, n, ret, weights=(1.0, 2.8)):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gosnark_b200 import shard as _shard
    from gosnark_b200.shard import shard_ranges
    _shard.W_AB, _shard.W_G2 = weights          # partition weights (B200_CFG_SHARD_W_AB / _W_G2 mirror)
    rng = random.Random(7)                       # same data on every rank (replicated inputs)
    m, npublic, n_ptd = n + 2, 1, n + 1
    ka, kb, kc, kp = ([rng.randrange(R) for _ in range(k)] for k in (m, m, m, n_ptd))
    w = [rng.randrange(R) for _ in range(m)]
    h = [rng.randrange(R) for _ in range(n - 1)]
    alpha, beta, delta, r, s = (rng.randrange(R) for _ in range(5))
    sh = shard_ranges(m, npublic, n_ptd, rank, world)
    S = sh["sets"]
    ncf = sh["n_c_full"]
    dot = lambda ks, ws, a, b: sum(ks[i] * ws[i] for i in range(a, b)) % R
    hpad = h + [0, 0]
    c_lo, c_hi = min(S[3]["lo"], ncf), min(S[3]["hi"], ncf)
    p_lo, p_hi = max(S[3]["lo"], ncf) - ncf, max(S[3]["hi"], ncf) - ncf
    part = [
        (dot(ka, w, S[0]["lo"], S[0]["hi"]) + (alpha + r * delta if S[0]["tail"] else 0)) % R,            # A
        (dot(kb, w, S[1]["lo"], S[1]["hi"]) + (beta + s * delta if S[1]["tail"] else 0)) % R,             # B1
        (dot(kb, w, S[2]["lo"], S[2]["hi"]) + (beta + s * delta if S[2]["tail"] else 0)) % R,             # B2 (same logs)
        (dot(kc[npublic + 1:], w[npublic + 1:], c_lo, c_hi) + dot(kp, hpad, p_lo, min(p_hi, n - 1))
         - (r * s * delta if S[3]["tail"] else 0)) % R,                                                    # C || PTD
    ]
    # 32-byte little-endian limbs in an int64 tensor, like the real partial records
    buf = torch.tensor([int.from_bytes(v.to_bytes(32, "little")[8 * k:8 * k + 8], "little", signed=False) - (1 << 63)
                        for v in part for k in range(4)], dtype=torch.int64)
    gathered = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    tot = [0, 0, 0, 0]
    for g in gathered:
        vals = [int(x) + (1 << 63) for x in g.tolist()]
        for j in range(4):
            tot[j] = (tot[j] + sum(vals[4 * j + k] << (64 * k) for k in range(4))) % R
    a_full = (dot(ka, w, 0, m) + alpha + r * delta) % R
    b_full = (dot(kb, w, 0, m) + beta + s * delta) % R
    c_full = (dot(kc, w, npublic + 1, m) + dot(kp, h, 0, n - 1) + s * a_full + r * b_full - r * s * delta) % R
    c_got = (tot[3] + s * tot[0] + r * tot[1]) % R      # groth16.go:272-275 applied after the gather
    ok = tot[0] == a_full and tot[1] == b_full and tot[2] == b_full and c_got == c_full
    # every element of every set is held by exactly one rank, every tail by exactly one rank
    for k, ln in enumerate((m, m, m, ncf + n_ptd)):
        cover = torch.zeros(ln + 1, dtype=torch.int64)
        cover[S[k]["lo"]:S[k]["hi"]] += 1
        cover[ln] += 1 if S[k]["tail"] else 0
        dist.all_reduce(cover)
        ok = ok and bool((cover == 1).all())
    if rank == 0:
        ret.put(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,weights", [(2, 64, (1.0, 2.8)), (3, 50, (1.0, 2.8)), (3, 50, (0.85, 2.75))])
def test_sharded_sums_over_gloo(world, n, weights):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(rk, world, port, n, ret, weights)) for rk in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
