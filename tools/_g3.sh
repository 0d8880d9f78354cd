tools/micro/pipe_mix > gpurun_out/r2_pipe_mix.txt 2>&1; cat gpurun_out/r2_pipe_mix.txt
timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -4
show() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$1', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],2), 'one', round(d['one_at_a_time']['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d.get('parity_vs_known_dlog'))"; }
timeout 400 python bench.py --no-extras > gpurun_out/r2_l_prove.json 2> gpurun_out/r2_l_prove.err; show gpurun_out/r2_l_prove.json
timeout 400 python bench.py --no-extras --logn 16 > gpurun_out/r2_l_prove16.json 2> gpurun_out/r2_l_prove16.err; show gpurun_out/r2_l_prove16.json
timeout 400 python bench.py --no-extras --workload g1msm > gpurun_out/r2_l_g1msm.json 2> gpurun_out/r2_l_g1msm.err; show gpurun_out/r2_l_g1msm.json
timeout 500 python bench.py --no-extras --workload g2msm > gpurun_out/r2_l_g2msm.json 2> gpurun_out/r2_l_g2msm.err; show gpurun_out/r2_l_g2msm.json
tail -3 gpurun_out/r2_l_*.err
