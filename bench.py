#!/usr/bin/env python
"""bench.py -- scenes/sec of the 3D-SIS dense-voxel TEST forward on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path (back-projection -> 3D backbone -> RPN/NMS -> RoI pool + classifier
-> per-RoI mask head) over one synthetic 96x48x96 ScanNet-shape chunk with 5 views
(BASELINE.json configs[1]); ENet-shaped 2D features are an input (ENet is upstream of the path,
SURVEY 8f2).  One process per GPU; chunks are independent so ranks shard with no data-path
collective ("weak" scaling: K chunks per rank); NCCL carries only the barrier and the max-over-ranks
time.  `value` times the forward with inputs resident in HBM, `e2e` the same call from pinned HOST
buffers including the result read-back.  See the task contract in DESIGN.md "measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "3d-sis_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

DIMS, N_IMG = (96, 48, 96), 5
ALG_BYTES = 806.8e6       # per-layer compulsory fp32 bytes, backbone + RPN, cfg2
MASK_BYTES_PER_VOXEL = 2712.0
METRIC = "scenes_per_sec_96x48x96_5img"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


def usable_cores():
    """Host threads this process may really use: affinity mask and cgroup CPU quota, not the box's core count."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return n


def best_threads(port, cores):
    """Pick the torch thread count that is fastest for the oracle on this host (over-subscription on a
    many-core box makes intra-op threading slower, and the baseline should be the best the CPU can do)."""
    import sis3d_synth as synth
    cfg = port.make_cfg("scannet", USE_IMAGES=False, USE_MASK=False)
    w = synth.make_weights(seed=0, use_images=False, use_mask=False)
    data, _ = synth.make_scene(7, (48, 32, 48))
    best, best_t = cores, float("inf")
    cand = sorted({c for c in (4, 8, 16, 32, 64, cores) if c <= cores})
    for c in cand:
        torch.set_num_threads(c)
        port.forward(cfg, w, data, None)
        t0 = time.perf_counter()
        port.forward(cfg, w, data, None)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def case(seed):
    import sis3d_synth as synth
    data, boxes = synth.make_scene(seed, DIMS)
    views = synth.make_views(seed, DIMS, N_IMG, boxes)
    return data, views


def weights():
    import sis3d_synth as synth
    return synth.make_weights(seed=0)


class Clocks:
    """nvidia-smi clock / throttle sampling during the timed region (one long-lived `-lms 100` process)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.stop_flag = index, None, False

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def summary(self):
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                rows = [[t.strip() for t in ln.split(",")] for ln in out.splitlines() if ln.strip()]
            except Exception:
                pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(rows)}


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference
    Python files cannot travel to the GPU box), all host threads, rank 0 only."""
    if rank != 0:
        return
    from oracle import port
    cores = best_threads(port, usable_cores())
    cfg = port.make_cfg("scannet")
    w = weights()
    data, views = case(303)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        port.forward(cfg, w, data, views)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
        if sum(times) > 150:  # bounded sample
            break
    ms = 1e3 * float(np.mean(times))
    v = 1e3 / ms
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "scenes/s", "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "96x48x96 ScanNet-shape chunk, 5 views, full rpn_class_mask_5 TEST forward (cfg2)",
                   "weights": "seeded synthetic", "threads": cores},
        "cpu_baseline": {"value": v, "unit": "scenes/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full chunks, torch-CPU fp32 oracle port of the reference path"},
        "e2e": {"value": v, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--chunks-per-step", type=int, default=32,
                    help="a step = one pass of the scene loop over this many chunks per rank (the pipeline holds 6 in flight)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sis3d")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true", help="profiling runs: skip the latency pass and the CPU baseline")
    ap.add_argument("--host-profile", type=int, default=0, help="cProfile N forwards -> gpurun_out/host_profile.txt")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from lib import _sis3d as S
    from test_gpu_forward import make_net
    from test_oracle_golden import CASES
    math = os.environ.get("SIS3D_CONV_MATH", "mixed")
    net, cfg = make_net(CASES["cfg2_96x48x96"], keep_debug=False, math=math)

    # 24 distinct chunks per rank rotate (24 x 6.9 MB = 166 MB > the 126 MB L2): host (pinned) and device copies
    n_in = 24
    host_in, dev_in = [], []
    for j in range(n_in):
        data, views = case(1000 + rank * 64 + j)
        hb = {"data": torch.from_numpy(data).pin_memory(), "id": ["bench"],
              "nearest_images": {k2: [torch.from_numpy(views[k1]).pin_memory()] for k1, k2 in
                                 (("feats", "images"), ("depths", "depths"), ("poses", "poses"), ("world2grid", "world2grid"))}}
        host_in.append(hb)
        dev_in.append({"data": hb["data"].to(dev), "id": ["bench"],
                       "nearest_images": {"images": [hb["nearest_images"]["images"][0].to(dev)],
                                          "depths": [hb["nearest_images"]["depths"][0].to(dev)],
                                          "poses": hb["nearest_images"]["poses"], "world2grid": hb["nearest_images"]["world2grid"]}})
    h2d = sum(t.numel() * t.element_size() for t in [host_in[0]["data"]] + [v[0] for v in host_in[0]["nearest_images"].values()])

    def step(blobs):  # one synchronous forward (latency view / per-kernel timing pass)
        return net.forward(blobs, "TEST", None)

    def timed_loop(inputs, steps):
        """K scenes through the scene-loop API (Network.forward_pipelined, 6 scenes in flight); every scene's
        detections and thresholded predicted-class masks are read back to the host.  CUDA events on the default
        stream bracket the region (it waits for the slot streams), barrier + synchronize on both sides."""
        d2h, vox, nroi, nmask = 0, 0, 0, 0
        k0 = net.kernel_launches()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dets = []
        for _, P in net.forward_pipelined(inputs[i % n_in] for i in range(steps)):
            det = P["detections_host"]  # results of this scene are on the host: detection table + thresholded masks
            d2h += det.nbytes + (P["mask_bits_host"].nbytes if "mask_bits_host" in P else 0)
            dets.append(det)
        for sl in net._slots:
            if sl["stream"] is not None:
                torch.cuda.current_stream().wait_stream(sl["stream"])
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for det in dets:  # workload statistics, outside the timed region
            k = det[det[:, 8] > 0.5]
            vox += int(((k[:, 12] - k[:, 9]) * (k[:, 13] - k[:, 10]) * (k[:, 14] - k[:, 11])).sum())
            nroi += det.shape[0]
            nmask += k.shape[0]
        return float(t.item()), net.kernel_launches() - k0, d2h / steps, vox / steps, nroi / steps, nmask / steps

    def latency(inputs, steps):
        ts = []
        for i in range(steps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(inputs[i % n_in])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    for i in range(args.warmup):
        step(dev_in[i % n_in])
    B = max(1, args.chunks_per_step)
    timed_loop(dev_in, max(8, args.warmup * B))   # warm-up steps; also captures the graphs of the pipeline slots
    timed_loop(host_in, max(8, args.warmup * B))
    if args.host_profile and rank == 0:
        import cProfile
        import io
        import pstats
        pr = cProfile.Profile()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pr.enable()
        timed_loop(host_in, args.host_profile)  # the pipelined scene loop, host buffers
        pr.disable()
        wall = (time.perf_counter() - t0) / args.host_profile * 1e3
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(40)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "host_profile.txt"), "w") as f:
            f.write(f"wall ms per scene (pipelined loop, host inputs, under cProfile): {wall:.3f}\n" + buf.getvalue())
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    n_chunks = args.steps * B  # chunks per rank inside each timed region
    ms_dev, launches, _, vox, nroi, nmask = timed_loop(dev_in, n_chunks)
    ms_e2e, _, d2h, _, _, _ = timed_loop(host_in, n_chunks)
    clocks.stop_flag = True
    # pinned-host -> device copy bandwidth of this box (explains the gap between `value` and `e2e`)
    probe_h = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
    probe_d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    probe_d.copy_(probe_h, non_blocking=True)
    torch.cuda.synchronize()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(8):
        probe_d.copy_(probe_h, non_blocking=True)
    pe1.record()
    torch.cuda.synchronize()
    h2d_gbs = 8 * (64 << 20) / (pe0.elapsed_time(pe1) * 1e-3) / 1e9
    del probe_h, probe_d
    lat_dev = latency(dev_in, 30) if not args.lean else None
    lat_host = latency(host_in, 30) if not args.lean else None

    # dominant tensor-core kernel, timed live: rpn_net_level{1,2} = 3x3x3 conv 128 -> 256 on the 24x12x24 level-1 grid through
    # the C ABI, 40 back-to-back launches between two CUDA events on the launching stream (inputs 3.5 MB: L2-resident, as
    # in the forward where the producer has just written them)
    from lib import _sis3d as S
    import ctypes as C
    rx = torch.randn(24, 12, 24, 128, device=dev)
    rw = torch.randn(256, 27 * 128, device=dev) * 0.02
    rb = torch.zeros(256, device=dev)
    ro = torch.empty(24, 12, 24, 256, device=dev)
    sh = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def rpn_conv():
        S.check(S.lib.sis3d_conv3d_k3_tc(S.ptr(rx), S.ptr(rw), S.ptr(rb), None, 0, 0, S.ptr(ro), 256, 0, 24, 12, 24, 128, 256, 3,
                                         None, 0, 1, sh), "rpn conv")
    for _ in range(5):
        rpn_conv()
    ke0, ke1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ke0.record()
    for _ in range(40):
        rpn_conv()
    ke1.record()
    torch.cuda.synchronize()
    rpn_ms = ke0.elapsed_time(ke1) / 40
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tflops, which = peaks()
    per_step_ms = ms_dev / args.steps
    per_chunk_ms = ms_dev / n_chunks
    value = world * n_chunks / (ms_dev / 1e3)
    e2e_v = world * n_chunks / (ms_e2e / 1e3)
    alg = ALG_BYTES + MASK_BYTES_PER_VOXEL * vox
    tf = 12.231e9 / (rpn_ms * 1e-3) / 1e12
    tensor_roof = {"bound": "tensor", "kernel": "conv3d_k3_tc_kernel<128,3> (rpn_net_level1/2: 3x3x3, 128 -> 256 ch, 24x12x24; TF32 in, fp32 accumulate)",
                   "achieved": tf, "peak": tflops / 2.0, "unit": "TFLOP/s", "frac": tf / (tflops / 2.0),
                   "peak_note": "TF32 dense = half of the measured bf16 GEMM peak in MEASURED_PEAKS.json (no TF32 figure measured)",
                   "flops_per_launch": 12.231e9, "ms_per_launch": rpn_ms, "launches_timed": 40,
                   "traffic": 7.1e6, "traffic_note": "dram__bytes_read+write per launch from the ncu --set full capture in profiles/ "
                                                     "(operands are L2 hits: 271 MB cross the L2->SM crossbar per launch)"}
    out = {
        "metric": METRIC, "value": value, "unit": "scenes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"tf32": "tf32 (convs on tcgen05, fp32 accumulate) + f32",
                  "mixed": "tf32 (static-stage convs) / f16 operands (mask-stage convs) on tcgen05, fp32 accumulate, + f32",
                  "fp16": "f16 operands (convs on tcgen05, fp32 accumulate) + f32"}.get(math, "f32"),
        "data": "synthetic",
        "config": {"workload": "96x48x96 ScanNet-shape chunk, 5 views, full rpn_class_mask_5 TEST forward (cfg2)",
                   "conv_math": math, "inputs": "seeded synthetic TSDF + ENet-shaped features/depth/poses; seeded synthetic weights",
                   "l2": "24 distinct chunks per rank rotate: 166 MB of inputs > 126 MB L2 (no flush kernel)",
                   "api": "Network.forward_pipelined (the scene loop; 6 scenes in flight on 6 streams: inputs uploaded one scene ahead, 3 graph replays overlapping)", "chunks_per_step": B,
                   "step": f"one pass of the scene loop over {B} chunks per rank", "rois_per_chunk": nroi,
                   "mask_rois_per_chunk": nmask, "mask_voxels_per_chunk": vox, "chunks_per_rank": n_chunks,
                   "parallelism": f"chunk-sharded dp{world}"},
        "e2e": {"value": e2e_v, "unit": "scenes/s", "h2d_bytes_per_step": h2d * B, "d2h_bytes_per_step": d2h * B,
                "ms_per_step": ms_e2e / args.steps, "ms_per_chunk": ms_e2e / n_chunks, "h2d_probe_gbs": round(h2d_gbs, 1),
                "what": "same loop from pinned HOST buffers: H2D of scene+features+depth+poses and D2H of detections + "
                        "thresholded predicted-class masks inside the timed region"},
        "latency_ms": {"sync_forward_device_inputs": lat_dev, "sync_forward_host_inputs": lat_host,
                       "note": "median of single synchronous Network.forward calls (no overlap between scenes)"},
        "gpu_launches": launches,
        # fraction of the 3D-conv HBM roofline = ALG_BYTES x scenes/s per GPU / HBM peak.  ALG_BYTES are the
        # per-layer compulsory fp32 bytes of the reference's dataflow; the fused/sparse design moves far fewer.
        "roofline": {"bound": "hbm", "achieved": alg * (value / world) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": alg * (value / world) / 1e9 / hbm, "traffic": 239.2e6, "peak_source": which,
                     "traffic_note": "dram__bytes_read+write summed over the launches of one scene in an ncu --set full capture "
                                     "(profiles/r1_ncu_full_one_scene_midround.txt; caches flushed per kernel -> upper bound; "
                                     "mid-round build)",
                     "kernel": "whole forward = all libsis3d launches of one scene (graph replay + ragged mask stage)",
                     "algorithmic_bytes_per_chunk": alg, "gpu_ms_per_chunk": per_chunk_ms,
                     "note": "per-kernel times and shares: profiles/ (ncu launch list + --set full capture of the same command)"},
        "roofline_tensor_kernel": tensor_roof,
        "clocks": clocks.summary(),
    }
    if not (args.no_cpu_baseline or args.lean) and world == 1:
        from oracle import port
        cores = best_threads(port, usable_cores())
        ocfg, w = port.make_cfg("scannet"), weights()
        data, views = case(1000)
        t_all, n = [], 0
        port.forward(ocfg, w, data, views)  # warm
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < args.cpu_baseline_seconds:
            t1 = time.perf_counter()
            port.forward(ocfg, w, data, views)
            t_all.append(time.perf_counter() - t1)
            n += 1
        out["cpu_baseline"] = {"value": 1.0 / float(np.mean(t_all)), "unit": "scenes/s", "cores": cores, "kind": "port",
                               "sample": f"{n} full cfg2 chunks in ~{args.cpu_baseline_seconds:.0f}s, torch-CPU fp32 oracle port"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
