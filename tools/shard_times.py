#!/usr/bin/env python3
"""Emulate every rank of an N-way sharded proof on ONE GPU and time each rank's partial prove (development aid:
shows which rank bounds the multi-GPU step)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from gosnark_b200 import _lib
from gosnark_b200._lib import check, ints_to_limbs, lib, ptr
from gosnark_b200.shard import shard_ranges
from gosnark_b200.synthetic import SyntheticGroth16

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
_lib.init(0)
L = lib()
# optional tuning: key=value pairs for b200_config (10 w_ab x100, 11 w_g2 x100, 12 affine min G1 terms, 13 affine min G2 terms)
from gosnark_b200 import shard as _shard
FLY = 1     # fly=2: two proofs in flight per emulated rank (two key contexts, two streams), as bench.py's default
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    if k == "fly":
        FLY = int(v)
        continue
    check(L.b200_config(int(k), int(v)))
    if int(k) == 10:
        _shard.W_AB = int(v) / 100.0
    if int(k) == 11:
        _shard.W_G2 = int(v) / 100.0
syn = SyntheticGroth16(logn)
r_l, s_l = ints_to_limbs([syn.r]), ints_to_limbs([syn.s])
d_w = torch.from_numpy(syn.w.view(np.int64)).cuda()
d_px = torch.from_numpy(syn.px.view(np.int64)).cuda()
d_out = torch.zeros(128, dtype=torch.int64, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
streams = [stream] + [torch.cuda.Stream() for _ in range(FLY - 1)]
d_outs = [d_out] + [torch.zeros(128, dtype=torch.int64, device="cuda") for _ in range(FLY - 1)]
for rk in range(world):
    pks = []
    for k in range(FLY):
        check(L.b200_config(_lib.CFG_PK_CONTEXT, k))
        pks.append(syn.load_pk(rk, world))
    check(L.b200_config(_lib.CFG_PK_CONTEXT, 0))
    pk = pks[0]

    def prove(i):
        k = i % FLY
        check(L.b200_groth16_prove_device(pks[k], d_w.data_ptr(), syn.m, d_px.data_ptr(), 2 * syn.n - 1, ptr(r_l), ptr(s_l), d_outs[k].data_ptr(), streams[k].cuda_stream))
    for i in range(3 * FLY):
        prove(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s_ in streams[1:]:
        s_.wait_event(e0)
    NP = 6
    for i in range(NP):
        prove(i)
    for s_ in streams[1:]:
        ev = torch.cuda.Event()
        ev.record(s_)
        stream.wait_event(ev)
    e1.record(stream)
    torch.cuda.synchronize()
    sh = shard_ranges(syn.m, syn.npublic, syn.n_ptd, rk, world)
    desc = " ".join(f"{n}[{s['lo']}:{s['hi']}{'+t' if s['tail'] else ''}]" for n, s in zip("A B1 B2 CH".split(), sh["sets"]) if s["lo"] < s["hi"] or s["tail"])
    ms = e0.elapsed_time(e1) / NP
    times = globals().setdefault("times", [])
    times.append(ms)
    print(f"rank {rk}/{world}: {ms:.3f} ms   {desc}")
    for pk_ in pks:
        check(L.b200_pk_free(pk_))
print(f"config {sys.argv[3:]}: slowest rank {max(times):.3f} ms, mean {sum(times)/len(times):.3f} ms")
