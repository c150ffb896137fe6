set -x
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12) > gpurun_out/r2p_test.log 2>&1
timeout 120 python __graft_entry__.py smoke
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/r2p_ncu_bench.log 2>&1
tail -5 gpurun_out/r2p_test.log
