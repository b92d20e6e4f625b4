timeout 1500 python tools/sc3_bench.py --cells 20000 --genes 20000 --no-cpu --single --out gpurun_out/sc3_bench20k.json > gpurun_out/sc3_bench20k.log 2>&1; echo rc=$? >> gpurun_out/sc3_bench20k.log
