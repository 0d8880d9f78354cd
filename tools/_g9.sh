python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err; python tools/show_bench.py gpurun_out/r2_bench_final2.json
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_final2.json') if l.startswith('{')][-1]); print(json.dumps(d.get('g1_msm_2p20'))); print(json.dumps(d.get('one_at_a_time'))[:200]); print(d['e2e']['ms_per_step'], d['e2e']['one_at_a_time_ms_per_step'])"
tail -n 3 gpurun_out/r2_bench_final2.err
