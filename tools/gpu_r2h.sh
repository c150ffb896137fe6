set -x
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30) > gpurun_out/r2h_test.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --lean > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
SIS3D_CLUSTER=1 timeout 300 python bench.py --steps 10 --warmup 3 --lean > gpurun_out/r2h_bench_cluster.json 2> gpurun_out/r2h_bench_cluster.err
timeout 300 python tools/parity_rate.py --modes exact > gpurun_out/r2h_parity.json 2> gpurun_out/r2h_parity.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/r2h_ncu_bench.log 2>&1
tail -6 gpurun_out/r2h_test.log; tail -c 400 gpurun_out/r2h_bench.err; tail -1 gpurun_out/r2h_parity.err
