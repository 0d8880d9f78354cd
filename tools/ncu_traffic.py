#!/usr/bin/env python3
"""Regenerate profiles/traffic.json (measured DRAM bytes per G1 term of the bucket-accumulation phase) from an ncu launch
list of bench.py taken with --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum.
usage: ncu_traffic.py launches.csv proofs g1_terms_per_proof"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_launch_summary import family, load

rows = load(sys.argv[1])
proofs, terms = int(sys.argv[2]), int(sys.argv[3])
tot = 0.0
for r in rows:
    f = family(r["name"])
    if f in ("k_affine_forward<Fq>", "k_affine_invert<Fq>", "k_affine_backward<Fq>", "k_accumulate<Fq>"):
        tot += r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
out = {"g1_accumulation_dram_bytes_per_term": tot / (proofs * terms),
       "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum summed over every G1 k_affine_forward / k_affine_invert / k_affine_backward "
                 f"launch of {proofs} proofs ({terms} G1 terms each: A, B1, C||PTD sets of a 2^20-constraint key; c = 17, 15 windows, S = 32): "
                 f"{tot / 1e6 / proofs:.0f} MB per proof; {os.path.basename(sys.argv[1])} (this round's final build, the launch list of "
                 "`bench.py --steps 2 --warmup 1 --no-extras --profile-region`)"}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
