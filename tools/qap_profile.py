#!/usr/bin/env python3
"""Development aid: time b200_qap_px (witness + sparse R1CS -> px) alone and bracket one call with cudaProfilerStart/Stop
(ncu --profile-from-start off).  usage: qap_profile.py [logn]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from gosnark_b200 import _lib  # noqa: E402
from gosnark_b200._lib import ints_to_limbs  # noqa: E402
from gosnark_b200.r1csqap import SparseR1CS  # noqa: E402
from gosnark_b200.synthetic import SyntheticCircuit  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
_lib.init(0)
c = SyntheticCircuit(1 << logn)
sp = SparseR1CS(c.n, c.m, c.csr)
w = ints_to_limbs(c.witness)
sp.combine_limbs(w, want_abc=False)          # builds the domain (one-time)
t0 = time.perf_counter()
for _ in range(3):
    sp.combine_limbs(w, want_abc=False)
print(f"qap_px 2^{logn}: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per call (host call incl. H2D of w, D2H of px)")
torch.cuda.profiler.start()
sp.combine_limbs(w, want_abc=False)
torch.cuda.profiler.stop()
