python -m pytest tests/test_gpu_nmf.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_gpu.log
python bench.py --cells 12500 --steps 50 --warmup 5 --no-e2e --no-cpu --sc3-cells 0 > gpurun_out/plain_shard2.log 2>&1
python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu --sc3-cells 0 > gpurun_out/plain_full2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wtx_|update|gram_|reduce_|iter_end" -s 30 -c 48 --csv --log-file gpurun_out/launches_shard2.csv python bench.py --cells 12500 --steps 20 --warmup 3 --no-e2e --no-cpu --sc3-cells 0 > gpurun_out/ncu_shard2.log 2>&1
