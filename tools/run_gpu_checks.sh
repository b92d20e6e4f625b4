python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t_gpu.log
python tools/kbench.py --k 8 32 50 --tc auto --out gpurun_out/kbench_upd2.json > gpurun_out/kbench_upd2.log 2>&1
