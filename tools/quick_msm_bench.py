#!/usr/bin/env python3
"""Quick device-resident MSM timing (development aid; bench.py is the contract)."""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gosnark_b200 import _lib, bn128  # noqa: E402
from oracle import ref_py as o  # noqa: E402


def rand_scalars(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 61) - 1)       # < 2^253 < r
    return a


def main():
    group = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    logn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    c = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    n = 1 << logn
    _lib.init(0)
    L = _lib.lib()
    words = 12 if group == 1 else 24
    ks = rand_scalars(n, 4)
    gen = bn128._flatten_g1([bn128.G1.G]) if group == 1 else bn128._flatten_g2([bn128.G2.G])
    pts = np.zeros((n, words), dtype=np.uint64)
    t0 = time.time()
    fn = L.b200_g1_mul_batch_bcast if group == 1 else L.b200_g2_mul_batch_bcast
    _lib.check(fn(_lib.ptr(gen), _lib.ptr(ks), n, _lib.ptr(pts)))
    print(f"mint {n} points: {time.time()-t0:.2f}s")
    t0 = time.time()
    h = _lib._h(0)
    fn = L.b200_g1_bases_load if group == 1 else L.b200_g2_bases_load
    _lib.check(fn(_lib.ptr(pts), n, c, h))
    print(f"bases_load: {time.time()-t0:.2f}s")
    bs = bn128.BaseSet.__new__(bn128.BaseSet)
    bs.group, bs.n, bs.handle = group, n, h.value
    print(bs.info())
    ss = rand_scalars(n, 5)
    # correctness via known dlogs
    t0 = time.time()
    out = bs.msm(limbs=ss)
    print(f"host msm call: {time.time()-t0:.4f}s")
    kk = _lib.limbs_to_ints(ks)
    sv = _lib.limbs_to_ints(ss)
    G = o.BN.G1 if group == 1 else o.BN.G2
    exp = G.affine(G.mul_scalar(G.G, sum(a * b for a, b in zip(kk, sv)) % o.R))
    ok = (out[0], out[1]) == (exp[0], exp[1])
    print("parity vs known-dlog expectation:", ok)
    # device-resident timing
    d_s = torch.from_numpy(ss.view(np.int64)).cuda()
    d_out = torch.zeros(64, dtype=torch.int64, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    for _ in range(3):
        _lib.check(L.b200_msm_device(bs.handle, d_s.data_ptr(), n, 0, d_out.data_ptr(), st))
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(L.b200_msm_device(bs.handle, d_s.data_ptr(), n, 0, d_out.data_ptr(), st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"G{group} MSM n=2^{logn}: {ms:.3f} ms  -> {n/ms/1e3:.1f} Mscalar-mul/s  ok={ok}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
