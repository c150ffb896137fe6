set -x
mkdir -p gpurun_out
nvidia-smi -L
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$R --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --lean > gpurun_out/r2j_cfg2_n2.json 2> gpurun_out/r2j_cfg2_n2.err
$R --master-port 29512 bench.py --gpus 2 --config cfg4 --steps 3 --warmup 3 --lean > gpurun_out/r2j_cfg4_n2.json 2> gpurun_out/r2j_cfg4_n2.err
$R --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2j_ref_n2.json 2> gpurun_out/r2j_ref_n2.err
tail -c 300 gpurun_out/r2j_cfg2_n2.err; tail -c 600 gpurun_out/r2j_cfg4_n2.err; tail -c 300 gpurun_out/r2j_cfg4_n2.json
