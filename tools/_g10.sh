timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_shard.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3
show() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$1', 'ms/step', round(d['ms_per_step'],3), 'one', round(d['one_at_a_time']['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d.get('parity_vs_known_dlog'))"; }
timeout 400 python bench.py --no-extras > gpurun_out/r2_s_prove.json 2> gpurun_out/r2_s_prove.err; show gpurun_out/r2_s_prove.json
timeout 400 python bench.py --no-extras --workload g1msm > gpurun_out/r2_s_g1msm.json 2> gpurun_out/r2_s_g1msm.err; show gpurun_out/r2_s_g1msm.json
timeout 400 python bench.py --no-extras --logn 16 > gpurun_out/r2_s_prove16.json 2> gpurun_out/r2_s_prove16.err; show gpurun_out/r2_s_prove16.json
