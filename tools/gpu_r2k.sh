set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$R --master-port 29521 bench.py --gpus 8 --config cfg4 --steps 5 --warmup 3 --lean > gpurun_out/r2k_cfg4_n8.json 2> gpurun_out/r2k_cfg4_n8.err
$R --master-port 29522 bench.py --gpus 8 --config cfg5 --steps 10 --warmup 3 --lean > gpurun_out/r2k_cfg5_n8.json 2> gpurun_out/r2k_cfg5_n8.err
$R --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 --lean > gpurun_out/r2k_cfg2_n8.json 2> gpurun_out/r2k_cfg2_n8.err
tail -c 300 gpurun_out/r2k_cfg4_n8.err; tail -c 300 gpurun_out/r2k_cfg5_n8.err
