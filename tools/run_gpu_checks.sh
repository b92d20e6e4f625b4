python -m pytest tests/test_gpu_nmf.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t_gpu.log
python tools/kbench.py --k 8 32 50 --tc auto --out gpurun_out/kbench_upd.json > gpurun_out/kbench_upd.log 2>&1
python bench.py --k 32 --steps 3 --warmup 3 --no-e2e > gpurun_out/plain_k32.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wtx_tc -s 4 -c 2 -o gpurun_out/prof_tc_k32_r1 python bench.py --k 32 --steps 3 --warmup 3 --no-e2e > gpurun_out/ncu_tc.log 2>&1
