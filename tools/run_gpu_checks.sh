python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t_gpu.log
python bench.py > gpurun_out/bench_s4.json 2> gpurun_out/bench_s4.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
