set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_enet.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout 200 python bench.py --enet --steps 5 --warmup 3 --lean > gpurun_out/r2o_enet.json 2> gpurun_out/r2o_enet.err
timeout 120 python tools/enet_check.py 2>&1 | tail -2
python -c "
import json; b=json.load(open('gpurun_out/r2o_enet.json')); print('enet path', round(b['value'],1), round(b['e2e']['value'],1))"
