timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -x -q -k "in_flight or two_host" 2>&1 | tail -4
show() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$1', 'ms/step', round(d['ms_per_step'],3), 'one', round(d['one_at_a_time']['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'e2e one', round(d['e2e']['one_at_a_time_ms_per_step'],3), d.get('parity_vs_known_dlog'))"; }
timeout 400 python bench.py --no-extras > gpurun_out/r2_n_prove.json 2> gpurun_out/r2_n_prove.err; show gpurun_out/r2_n_prove.json
timeout 400 python bench.py --no-extras --logn 16 > gpurun_out/r2_n_prove16.json 2> gpurun_out/r2_n_prove16.err; show gpurun_out/r2_n_prove16.json
tail -n 3 gpurun_out/r2_n_prove.err
