python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --sc3-cells 0 > gpurun_out/plain_final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wtx_|update_kernel|gram_|reduce_|iter_end|tc_shadow|shadows|transpose" -c 200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --sc3-cells 0 > gpurun_out/ncu_final1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wtx_tc -s 6 -c 2 -o gpurun_out/prof_tc_k8_r1 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --sc3-cells 0 > gpurun_out/ncu_final2.log 2>&1
timeout 900 python tools/sc3_bench.py --cells 20000 --genes 20000 --no-cpu --single --out gpurun_out/sc3_bench20k_v2.json > gpurun_out/sc3_bench20k_v2.log 2>&1
