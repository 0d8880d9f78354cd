# partition weights under two proofs in flight (bench.py's mode at N > 1): A / B1 weight (key 10) and G2 weight (key 11), x100
mkdir -p gpurun_out
: > gpurun_out/r2_shard_weights_fly2.log
run() { timeout 120 python tools/shard_times.py 20 $1 fly=2 12=1 13=1 10=$2 11=$3 2>&1 | tail -$(( $1 + 1 )) >> gpurun_out/r2_shard_weights_fly2.log; }
run 8 85 275
run 8 90 280
run 8 85 265
run 8 80 270
run 8 90 270
run 4 85 275
run 4 90 270
run 2 100 280
run 2 85 275
run 2 90 270
grep "^config" gpurun_out/r2_shard_weights_fly2.log
