set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/r2a_test.log 2>&1
(time python tools/parity_rate.py > gpurun_out/r2a_parity.json) 2> gpurun_out/r2a_parity.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_exact.json 2> gpurun_out/r2a_bench_exact.err
SIS3D_CONV_MATH=mixed python bench.py --steps 20 --warmup 3 --lean > gpurun_out/r2a_bench_mixed.json 2> gpurun_out/r2a_bench_mixed.err
python tools/loop_pipelined.py --nets 6 --reps 25 > gpurun_out/r2a_loop.json 2> gpurun_out/r2a_loop.err
python tools/enet_check.py > gpurun_out/r2a_enet.log 2>&1
timeout 300 compute-sanitizer --tool synccheck python __graft_entry__.py smoke > gpurun_out/r2a_synccheck_smoke.log 2>&1
timeout 400 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r2a_racecheck_smoke.log 2>&1
timeout 500 compute-sanitizer --tool racecheck python tools/loop_pipelined.py --nets 1 --reps 1 > gpurun_out/r2a_racecheck_pipelined.log 2>&1
tail -3 gpurun_out/r2a_test.log; cat gpurun_out/r2a_loop.json; tail -2 gpurun_out/r2a_enet.log; tail -3 gpurun_out/r2a_racecheck_smoke.log
