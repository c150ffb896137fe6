# gpurun -- bash tools/gpurun_measure.sh : the bench lines of every BASELINE config, the reference arm, the ncu launch list (with DRAM bytes)
# and the --set full captures of the top kernels that profiles/r2_* were taken from (round 2).
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/meas_bench_cfg2.json 2> gpurun_out/meas_bench_cfg2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/meas_bench_ref.json 2> gpurun_out/meas_bench_ref.err
python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/meas_bench_cfg5.json 2> gpurun_out/meas_bench_cfg5.err
python bench.py --config cfg3s --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/meas_bench_cfg3s.json 2> gpurun_out/meas_bench_cfg3s.err
python bench.py --config cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/meas_bench_cfg3.json 2> gpurun_out/meas_bench_cfg3.err
python bench.py --config cfg4 --steps 3 --warmup 3 --lean > gpurun_out/meas_bench_cfg4_n1.json 2> gpurun_out/meas_bench_cfg4_n1.err
python bench.py --enet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/meas_bench_cfg2_enet.json 2> gpurun_out/meas_bench_cfg2_enet.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/meas_launches.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/meas_ncu_bench.log 2>&1
i=0
for k in ".int.128, .int.3, .int.4, .int.128, .int.2, .int.0, .int.2, .int.1" ".int.32, .int.3, .int.4, .int.128, .int.2, .int.32, .int.2, .int.1" ".int.64, .int.3, .int.2, .int.128, .int.4, .int.0, .int.0, .int.1"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv3d_k3_tc_kernel<$k>" -s 30 -c 2 -o gpurun_out/meas_full_$i -f python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean > gpurun_out/meas_ncu_full_$i.log 2>&1
done
for f in cfg2 cfg5 cfg3s cfg3 cfg4_n1 cfg2_enet; do tail -c 300 gpurun_out/meas_bench_$f.err; done
