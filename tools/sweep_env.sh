#!/bin/bash
# usage: tools/sweep_env.sh "VAR=1 VAR2=2" "VAR=3" ...   -> one bench.py line (ms/step, e2e ms, parity, G1/G2 phase ms) per setting
for cfg in "$@"; do
  echo -n "[$cfg] "
  env $cfg python bench.py --steps 6 --warmup 3 --no-extras 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read())
r = d["roofline"]
print(round(d["ms_per_step"], 3), round(d["e2e"]["ms_per_step"], 3), d["parity_vs_known_dlog"], round(r["avg_launch_ms"], 3), round(r["g2"]["avg_launch_ms"], 3))'
done
