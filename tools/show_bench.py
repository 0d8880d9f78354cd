#!/usr/bin/env python3
"""Print the key numbers of a bench.py JSON line."""
import json
import sys
for path in sys.argv[1:]:
    lines = [l for l in open(path) if l.startswith("{")]
    if not lines:
        print(path, "no JSON line")
        continue
    d = json.loads(lines[-1])
    rf = d.get("roofline") or {}
    print(f"{path}: n_gpus={d.get('n_gpus')} ms/step={d.get('ms_per_step'):.3f} value={d.get('value'):.2f} "
          f"e2e_ms={d['e2e'].get('ms_per_step', 0):.3f} parity={d.get('parity_vs_known_dlog')} "
          f"launches={d.get('gpu_launches')} clocks={d.get('clocks')} "
          f"roofline_frac={rf.get('frac')} g1_phase_ms={rf.get('avg_launch_ms')}")
