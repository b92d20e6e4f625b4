python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t_gpu.log
python tools/kbench.py --k 8 16 32 50 --tc auto --out gpurun_out/kbench_tc_counts.json > gpurun_out/kbench_tc_counts.log 2>&1
python tools/kbench.py --k 8 16 32 50 --tc auto --real --out gpurun_out/kbench_tc_real.json > gpurun_out/kbench_tc_real.log 2>&1
python bench.py --no-cpu > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err
