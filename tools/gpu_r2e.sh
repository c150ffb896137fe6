set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30) > gpurun_out/r2e_test.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_exact.json 2> gpurun_out/r2e_bench_exact.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/r2e_ncu_bench.log 2>&1
i=0
for k in ".int.128, .int.3, .int.4, .int.128, .int.2, .int.0, .int.2" ".int.32, .int.3, .int.4, .int.128, .int.2, .int.32, .int.2" ".int.64, .int.3, .int.2, .int.128, .int.4, .int.0, .int.0" ".int.32, .int.1, .int.4, .int.128, .int.2, .int.0, .int.2"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv3d_k3_tc_kernel<$k>" -s 30 -c 2 -o gpurun_out/r2e_full_$i -f python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/r2e_ncu_full_$i.log 2>&1
done
tail -4 gpurun_out/r2e_test.log; ls -la gpurun_out/*.ncu-rep
