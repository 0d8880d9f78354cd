timeout 900 python -m pytest tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -3
for l in 16 12 4; do
timeout 300 python bench.py --workload verify --pairing-kernel 2 --logn $l --steps 3 --warmup 1 > gpurun_out/r2_r_verify_k2_$l.json 2> gpurun_out/r2_r_verify_k2_$l.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_r_verify_k2_$l.json') if l.startswith('{')][-1]); print('kernel 2 (128 regs) logn $l', round(d['value']), 'pairings/s', round(d['ms_per_step'],3), 'ms/batch', d['parity_bilinearity'], 'verify ms', round(d['groth16_verify']['ms_per_call'],3), d['groth16_verify']['accepted'])"
done
timeout 300 python bench.py --workload verify --pairing-kernel 1 --logn 12 --steps 3 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('kernel 1 logn 12', round(d['value']), 'pairings/s', round(d['ms_per_step'],3))"
