set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60) > gpurun_out/r2c_test.log 2>&1
python tools/parity_rate.py --modes exact tf32x3 > gpurun_out/r2c_parity.json 2> gpurun_out/r2c_parity.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench_exact.json 2> gpurun_out/r2c_bench_exact.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/r2c_ncu_bench.log 2>&1
python bench.py --config cfg5 --steps 5 --warmup 3 --lean > gpurun_out/r2c_bench_cfg5.json 2> gpurun_out/r2c_bench_cfg5.err
python bench.py --config cfg3s --steps 3 --warmup 3 --chunks-per-step 32 --lean > gpurun_out/r2c_bench_cfg3s.json 2> gpurun_out/r2c_bench_cfg3s.err
python bench.py --config cfg3 --steps 2 --warmup 3 --chunks-per-step 16 --lean > gpurun_out/r2c_bench_cfg3.json 2> gpurun_out/r2c_bench_cfg3.err
tail -5 gpurun_out/r2c_test.log; tail -2 gpurun_out/r2c_parity.err; tail -c 300 gpurun_out/r2c_bench_cfg3.err
