set -x
mkdir -p gpurun_out
for n in 4 6 8; do
  SIS3D_PIPE_STATIC=$n timeout 200 python bench.py --enet --steps 5 --warmup 3 --lean > gpurun_out/r2n_enet_static$n.json 2> gpurun_out/r2n_enet_static$n.err
done
python - <<'PY'
import json
for n in (4,6,8):
    try:
        b=json.load(open(f"gpurun_out/r2n_enet_static{n}.json")); print(n, round(b["value"],1), round(b["e2e"]["value"],1))
    except Exception as e:
        print(n, "ERR", e)
PY
