#!/bin/bash
# First GPU call of the next round: validate and time the prepared experiments.
#   gpurun --timeout 900 -- 'tools/r2_experiments.sh > gpurun_out/r2_experiments.log 2>&1; tail -40 gpurun_out/r2_experiments.log'
# 1. parity of the pipelined kernels (inlined and out-of-line multiply), 2. bench sweep, 3. the Karatsuba build variant.
set -u
cd "$(dirname "$0")/.."
echo "== parity, B200_AFF_SP=3 (inlined multiply, G1)"
B200_AFF_SP=3 B200_ACC_MODE=affine timeout 300 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q 2>&1 | tail -2
echo "== parity, B200_AFF_SP=7 (out-of-line multiply, G1 + G2)"
B200_AFF_SP=7 B200_ACC_MODE=affine timeout 300 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q 2>&1 | tail -2
echo "== parity, B200_AFF_LR=3 (short-live-range backward, G1 + G2)"
B200_AFF_LR=3 B200_ACC_MODE=affine timeout 300 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q 2>&1 | tail -2
echo "== bench sweep: ms/step, e2e ms, parity, G1 phase ms, G2 phase ms"
timeout 400 tools/sweep_env.sh "B200_AFF_LR=1" "B200_AFF_LR=2" "B200_AFF_LR=3" "B200_AFF_LR=3 B200_AFF_MINB_G2=5" \
  "B200_AFF_LR=3 B200_AFF_MINB=6 B200_AFF_MINB_G2=5" "B200_AFF_LR=3 B200_AFF_MINB=6 B200_AFF_MINB_G2=6"
timeout 500 tools/sweep_env.sh "B200_X=0" "B200_AFF_SP=1" "B200_AFF_SP=2" "B200_AFF_SP=3" "B200_AFF_SP=3 B200_AFF_MINB=3" \
  "B200_AFF_SP=5" "B200_AFF_SP=6" "B200_AFF_SP=7" "B200_AFF_SP=3 B200_AFF_MINB_FWD=6"
echo "== fewer pairs per thread in the late rounds (B200_AFF_TSMALL = CTA threshold): parity, bench, stand-alone MSMs"
B200_AFF_TSMALL=1184 B200_ACC_MODE=affine timeout 300 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q 2>&1 | tail -2
timeout 200 tools/sweep_env.sh "B200_AFF_TSMALL=600" "B200_AFF_TSMALL=1184" "B200_AFF_TSMALL=2400"
for ts in 0 1184; do echo -n "G1 2^20 stand-alone, TSMALL=$ts: "; B200_AFF_TSMALL=$ts timeout 100 python tools/quick_msm_bench.py 1 20 16 2>&1 | tail -1; done
echo "== thread-per-slice fused rounds (B200_AFF_TS: bit 0 G1, bit 1 G2): parity, bench, stand-alone MSMs"
B200_AFF_TS=3 B200_ACC_MODE=affine timeout 300 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_shard.py -x -q 2>&1 | tail -2
timeout 300 tools/sweep_env.sh "B200_AFF_TS=1" "B200_AFF_TS=2" "B200_AFF_TS=3" "B200_AFF_TS=3 B200_AFF_TS_MINB=3" "B200_AFF_TS=3 B200_AFF_TS_SMEM=1" "B200_AFF_TS=3 B200_AFF_TS_SMEM=1 B200_AFF_TS_MINB=5"
for g in 1 2; do echo -n "G$g 2^20 stand-alone, TS: "; B200_AFF_TS=3 timeout 100 python tools/quick_msm_bench.py $g 20 16 2>&1 | tail -1; done
echo "== fused NTT passes (B200_NTT_FUSED=1): parity of the polynomial and prove tests, then bench"
B200_NTT_FUSED=1 timeout 300 python -m pytest tests/test_gpu_poly.py tests/test_gpu_prove.py -x -q 2>&1 | tail -2
timeout 120 tools/sweep_env.sh "B200_NTT_FUSED=1"
echo "== fast final exponentiation: parity on the goldens, then timing of the verify tests"
B200_FAST_FINAL_EXP=1 timeout 200 python -m pytest tests/test_gpu_verify.py -x -q --durations=4 2>&1 | tail -8
if [ -f go-snark-study_b200/lib/libb200snark_k.so ]; then
  echo "== Karatsuba build variant: parity, then bench"
  B200_LIB_VARIANT=k timeout 300 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_poly.py -x -q 2>&1 | tail -2
  timeout 200 tools/sweep_env.sh "B200_LIB_VARIANT=k" "B200_LIB_VARIANT=k B200_AFF_SP=3"
else
  echo "(build the variant first, in the CPU container: python -c \"import build; build.build_cuda(variant='k')\")"
fi
