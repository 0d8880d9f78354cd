# N-rank bench lines after the px top-half staging (usage: bash tools/_g12.sh 2)
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$1', 'ms/step', round(d['ms_per_step'],3), 'one', round(d['one_at_a_time']['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'e2e one', round(d['e2e']['one_at_a_time_ms_per_step'],3), d.get('parity_vs_known_dlog'))"; }
for N in "$@"; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_px_scale_n$N.json 2> gpurun_out/r2_px_scale_n$N.err; show gpurun_out/r2_px_scale_n$N.json
tail -n 2 gpurun_out/r2_px_scale_n$N.err
done
