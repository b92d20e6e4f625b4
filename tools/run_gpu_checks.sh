python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t_full.log
