set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r2m_bench_cfg2.json 2> gpurun_out/r2m_bench_cfg2.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/r2m_ncu_bench.log 2>&1
timeout 120 python -m pytest tests/test_gpu_forward.py -m gpu -q -p no:cacheprovider -k "pipelined or graph" 2>&1 | tail -3
tail -c 500 gpurun_out/r2m_bench_cfg2.err
