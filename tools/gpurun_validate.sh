# gpurun -- bash tools/gpurun_validate.sh : GPU test suite, pipeline-depth A/B, compute-sanitizer synccheck / racecheck on smoke() and on the scene loop.
set -x
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12) > gpurun_out/val_test.log 2>&1
SIS3D_PIPE_STATIC=4 timeout 200 python bench.py --steps 10 --warmup 3 --lean > gpurun_out/val_bench_static4.json 2> gpurun_out/val_bench_static4.err
SIS3D_PIPE_STATIC=5 timeout 200 python bench.py --steps 10 --warmup 3 --lean > gpurun_out/val_bench_static5.json 2> gpurun_out/val_bench_static5.err
timeout 200 python bench.py --steps 10 --warmup 3 --lean > gpurun_out/val_bench_static3.json 2> gpurun_out/val_bench_static3.err
timeout 300 compute-sanitizer --tool synccheck python __graft_entry__.py smoke > gpurun_out/val_synccheck_smoke.log 2>&1
timeout 400 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/val_racecheck_smoke.log 2>&1
timeout 400 compute-sanitizer --tool racecheck python tools/loop_pipelined.py --nets 1 --reps 1 > gpurun_out/val_racecheck_pipelined.log 2>&1
tail -3 gpurun_out/val_test.log; tail -2 gpurun_out/val_racecheck_smoke.log; tail -2 gpurun_out/val_racecheck_pipelined.log; tail -2 gpurun_out/val_synccheck_smoke.log
