# closing check of the round: full GPU suite on the final library, then the slice-size step probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_closing.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_gputest_closing.log
timeout 60 python tools/slice_cliff.py 700000 832000 838000 845000 891000 > gpurun_out/r2_slice_cliff.log 2>&1; cat gpurun_out/r2_slice_cliff.log | tail -6
