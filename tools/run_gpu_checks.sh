python -m pytest tests/test_gpu_nmf.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t_gpu.log
python tools/init_bench.py > gpurun_out/init_bench.json 2> gpurun_out/init_bench.err
python tools/sc3_bench.py --cells 4000 --genes 4000 > gpurun_out/sc3_bench4k.log 2>&1; cp gpurun_out/sc3_bench.json gpurun_out/sc3_bench4k.json
