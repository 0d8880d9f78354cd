#!/usr/bin/env python3
"""Run a few device-resident Groth16 proofs (for ncu launch lists / profiles)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from gosnark_b200 import _lib
from gosnark_b200._lib import check, ints_to_limbs, lib, ptr
from gosnark_b200.synthetic import SyntheticGroth16

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
_lib.init(0)
L = lib()
syn = SyntheticGroth16(logn)
pk = syn.load_pk()
r_l, s_l = ints_to_limbs([syn.r]), ints_to_limbs([syn.s])
d_w = torch.from_numpy(syn.w.view(np.int64)).cuda()
d_px = torch.from_numpy(syn.px.view(np.int64)).cuda()
d_out = torch.zeros(48, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
for i in range(reps):
    if i == reps - 1:
        torch.cuda.nvtx.range_push("last")
    check(L.b200_groth16_prove_device(pk, d_w.data_ptr(), syn.m, d_px.data_ptr(), 2 * syn.n - 1, ptr(r_l), ptr(s_l), d_out.data_ptr(), None))
    torch.cuda.synchronize()
print("done")
