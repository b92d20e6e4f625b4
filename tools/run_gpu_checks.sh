python -m pytest tests/test_gpu_sc3.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t_gpu.log
timeout 600 python tools/svd_bench.py 4000 8000 20000 > gpurun_out/svd_plain.log 2>&1
