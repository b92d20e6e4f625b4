python bench.py --graph 1 --steps 50 --warmup 5 --no-cpu --sc3-cells 0 --no-e2e > gpurun_out/graph1.log 2>&1; tail -2 gpurun_out/graph1.log
python bench.py --graph 1 --cells 12500 --steps 50 --warmup 5 --no-cpu --sc3-cells 0 > gpurun_out/graph_shard.json 2> gpurun_out/graph_shard.err; tail -2 gpurun_out/graph_shard.err
