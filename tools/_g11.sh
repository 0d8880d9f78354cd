# round-2 closing check of the px top-half staging: full GPU suite, then the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_px.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_gputest_px.log
timeout 420 python bench.py > gpurun_out/r2_bench_n1_px.json 2> gpurun_out/r2_bench_n1_px.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_n1_px.json
