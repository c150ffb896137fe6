set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/r2d_test.log 2>&1
python tools/parity_rate.py --modes exact > gpurun_out/r2d_parity.json 2> gpurun_out/r2d_parity.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench_exact.json 2> gpurun_out/r2d_bench_exact.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/r2d_ncu_bench.log 2>&1
for k in "128, 3, 4, 128, 2, 0, 2" "32, 3, 4, 128, 2, 32, 2" "64, 3, 2, 128, 4, 0, 0"; do
  tag=$(echo $k | tr -d ' ,')
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv3d_k3_tc_kernel<$k>" -s 30 -c 2 -o gpurun_out/r2d_full_$tag -f python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/r2d_ncu_full_$tag.log 2>&1
done
tail -4 gpurun_out/r2d_test.log; tail -1 gpurun_out/r2d_parity.err; ls -la gpurun_out/*.ncu-rep
