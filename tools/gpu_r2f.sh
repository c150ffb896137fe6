set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30) > gpurun_out/r2f_test.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench_exact.json 2> gpurun_out/r2f_bench_exact.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/r2f_ncu_bench.log 2>&1
python tools/loop_pipelined.py --nets 4 --reps 25 > gpurun_out/r2f_loop.json 2> gpurun_out/r2f_loop.err
tail -4 gpurun_out/r2f_test.log; cat gpurun_out/r2f_loop.json
