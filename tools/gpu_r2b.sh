set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15) > gpurun_out/r2b_test.log 2>&1
python tools/parity_rate.py --modes exact > gpurun_out/r2b_parity.json 2> gpurun_out/r2b_parity.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench_exact.json 2> gpurun_out/r2b_bench_exact.err
SIS3D_PIPE_STATIC=4 python bench.py --steps 10 --warmup 3 --lean > gpurun_out/r2b_bench_exact_static4.json 2> gpurun_out/r2b_bench_exact_static4.err
SIS3D_PIPE_STATIC=2 python bench.py --steps 10 --warmup 3 --lean > gpurun_out/r2b_bench_exact_static2.json 2> gpurun_out/r2b_bench_exact_static2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/r2b_ncu_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
tail -3 gpurun_out/r2b_test.log; cat gpurun_out/r2b_parity.err | tail -2; tail -c 600 gpurun_out/r2b_bench_ref.json
