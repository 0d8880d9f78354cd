# GLV icPubl + phase-cost partition: full GPU suite, verify bench, 8- and 4-way shard emulation sweeps on one GPU
# (record of the command that produced profiles/r2_shard_phase_cost.log; config key 14 — the phase cost — was measured worse and
#  removed from the library afterwards, so this script no longer runs as is)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_part.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_gputest_part.log
timeout 200 python bench.py --workload verify > gpurun_out/r2_bench_verify_glv.json 2> gpurun_out/r2_bench_verify_glv.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_verify_glv.json') if l.startswith('{')][-1]); print('verify', round(d['value']), d['unit'], d['groth16_verify'])"
: > gpurun_out/r2_shard_phase_cost.log
for F in 0 80000 160000 240000; do
  timeout 200 python tools/shard_times.py 20 8 fly=2 12=1 13=1 14=$F 2>&1 | tail -9 >> gpurun_out/r2_shard_phase_cost.log
done
for F in 0 160000; do
  timeout 200 python tools/shard_times.py 20 4 fly=2 12=1 13=1 14=$F 2>&1 | tail -5 >> gpurun_out/r2_shard_phase_cost.log
done
grep "^config" gpurun_out/r2_shard_phase_cost.log
