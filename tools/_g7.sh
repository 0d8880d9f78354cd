timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gputest_final.log; tail -3 gpurun_out/r2_gputest_final.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-extras --profile-region > gpurun_out/r2_bench_under_ncu_final.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r2_launches_bench_final.csv > gpurun_out/r2_launches_bench_final_summary.txt; head -8 gpurun_out/r2_launches_bench_final_summary.txt
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; python tools/show_bench.py gpurun_out/r2_bench_final.json
timeout 300 python bench.py --workload verify > gpurun_out/r2_bench_verify_final.json 2> gpurun_out/r2_bench_verify_final.err; tail -c 400 gpurun_out/r2_bench_verify_final.json
