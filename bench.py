#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-ray sample pipeline (contract: task prompt ④).

    python bench.py --gpus N --steps K --warmup W            # product arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

N=1 workload = BASELINE.json configs[1]: NeRF-synthetic lego-shaped full-frame render 800x800,
synthetic 1M neural points (voxel query + aggregation + compositing).  metric: render Mpix/s.
One step = one full frame.  N>1: frames are independent units -> each rank renders its own frame
(weak scaling, no data-path collective); value = frames of all ranks / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 800
N_POINTS = 1_000_000
V = 4
CHUNK = None          # whole frame per query (one host readback per frame)
WORKLOAD = "NeRF-synthetic lego-shaped full-frame render 800x800, synthetic 1M neural points (voxel query + aggregation + compositing)"
CPU_SAMPLE_RAYS = 20480      # 160 x 128 window at the image centre: ~12 s of host work per pass on 16 threads
CPU_CHUNK_RAYS = 1024         # walked in chunks (the reference's own frame driver renders chunk by chunk, train_ft.py:282-351)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every 20 ms; nvidia-smi as a fallback)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self.run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else \
            n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        g = lambda name, alt: getattr(n, name, getattr(n, alt, 0))
        bits = [g("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                g("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                g("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                g("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap")]
        return [str(sm), str(mx)] + ["Active" if (b and (r & b)) else "Not Active" for b in bits]

    def run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml())
                    time.sleep(0.02)
                    continue
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                self.nvml = None
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def build_scene(seed):
    from hybridneuralrendering_b200 import synthetic as syn
    xyz = syn.lego_scene(N_POINTS, seed)
    att = syn.point_attributes(np.random.default_rng(seed), len(xyz))
    fr = syn.lego_frame(H=H, W=W, V=V, seed=seed)
    return xyz, att, fr


def build_net(xyz, att, dev, P):
    from hybridneuralrendering_b200 import NeuralPoints, NeuralPointsRayMarching, PointAggregator, make_opt
    opt = make_opt("lego", use_nearest=V, is_train=False)
    c = lambda a: torch.from_numpy(a).to(dev)
    pts = NeuralPoints(32, len(xyz), opt, dev)
    pts.set_points(c(xyz), c(att["emb"])[None], points_color=c(att["color"])[None], points_dir=c(att["dir"])[None],
                   points_conf=c(att["conf"])[None], parameter=True)
    agg = PointAggregator(opt).to(dev)
    agg.load_state_dict(P, strict=False)
    net = NeuralPointsRayMarching(aggregator=agg, neural_points=pts, opt=opt).to(dev)
    net.near_far = (2.0, 6.0)
    return net, opt


FRAME_KEYS = ("campos", "camrotc2w", "raydir", "near", "far", "intrinsic", "bg_color", "images_nearest", "c2w_nearest", "campos_nearest",
              "intrinsic_nearest")


def cpu_reference_sample(P, xyz, att, fr, opt, q_np, n_threads):
    """the reference's CPU path for aggregation + compositing (oracle port of PointAggregator.forward,
    the ray_dist prologue and ray_march), timed on a bounded sample of the frame's kept rays."""
    from oracle import pipeline_oracle as po
    from oracle import render_oracle as ro
    torch.set_num_threads(n_threads)
    cfg = ro.AggCfg(use_nearest=V)
    n = int(q_np["sample_pidx"].shape[1])
    pts = dict(xyz=xyz, **att)
    out = None
    t0 = time.perf_counter()
    with torch.no_grad():
        for r0 in range(0, n, CPU_CHUNK_RAYS):
            qc = {k: q_np[k][:, r0:r0 + CPU_CHUNK_RAYS] for k in ("sample_pidx", "sample_loc", "sample_loc_w", "sample_ray_dirs")}
            out = po.render_from_query(P, cfg, pts, qc, fr, float(opt.vsize[2]))
    dt = time.perf_counter() - t0
    return out, dt


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the Python
    reference cannot travel to the GPU box), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import render_oracle as ro
    dev = torch.device("cuda:0") if torch.cuda.is_available() else None
    xyz, att, fr = build_scene(0)
    P = ro.random_params(0)
    cores = os.cpu_count() or 1
    q_np, n_sample = sample_query(xyz, att, fr, P, dev)
    times = []
    for i in range(args.warmup + args.steps):
        _, dt = cpu_reference_sample(P, xyz, att, fr, make_opt_lego(), q_np, cores)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = n_sample / t / 1e6
    line = {"impl": "reference", "metric": "render Mpix/s", "value": val, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": f"{n_sample} kept rays of the frame (aggregation+compositing on CPU; query results precomputed)"},
            "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": f"{n_sample} kept rays, torch {torch.__version__} CPU fp32"},
            "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def make_opt_lego():
    from hybridneuralrendering_b200 import make_opt
    return make_opt("lego", use_nearest=V, is_train=False)


def sample_query(xyz, att, fr, P, dev):
    """query tensors (numpy) for a bounded sample of rays around the image centre.  The reference has no
    CPU query (its query is CUDA only), so the sample's neighbour lists come from the product query on the
    GPU when one is present, else from the numpy oracle on a smaller sample."""
    yy, xx = np.meshgrid(np.arange(H // 2 - 80, H // 2 + 80), np.arange(W // 2 - 64, W // 2 + 64), indexing="ij")
    ids = (yy * W + xx).reshape(-1)[:CPU_SAMPLE_RAYS]
    sub = dict(fr, raydir=fr["raydir"][:, ids])
    opt = make_opt_lego()
    if dev is not None:
        net, _ = build_net(xyz, att, dev, P)
        c = lambda a: torch.from_numpy(a).to(dev)
        inputs = {"raydir": c(sub["raydir"]), "campos": c(fr["campos"]), "camrotc2w": c(fr["camrotc2w"])}
        pidx, loc, loc_w, dirs, mask, _, _ = net.neural_points.query(inputs, near=2.0, far=6.0)
        q = dict(sample_pidx=pidx.cpu().numpy(), sample_loc=loc.cpu().numpy(), sample_loc_w=loc_w.cpu().numpy(), sample_ray_dirs=dirs.cpu().numpy())
        del net
        torch.cuda.empty_cache()
    else:
        from oracle import query_oracle as qo
        ids = ids[:64]
        sub = dict(fr, raydir=fr["raydir"][:, ids])
        ts = qo.candidate_ts(int(opt.z_depth_dim), 2.0, 6.0)[0, 0]
        q = qo.query(xyz, fr["campos"], fr["camrotc2w"], sub["raydir"], ts, vsize=opt.vsize, vscale=opt.vscale, kernel_size=opt.kernel_size,
                     query_size=opt.query_size, ranges=opt.ranges, radius_limit_scale=opt.radius_limit_scale, SR=opt.SR, K=opt.K, P=opt.P)
    return q, int(q["sample_pidx"].shape[1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step half of the metric")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch.distributed as dist
    from hybridneuralrendering_b200 import ops
    from hybridneuralrendering_b200.renderer import render_rays
    from oracle import render_oracle as ro          # only for the seeded weights + the cpu_baseline leg

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback of the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    xyz, att, fr = build_scene(rank)               # every rank renders its own view of its own replica
    P = ro.random_params(0)
    net, opt = build_net(xyz, att, dev, P)
    host = {k: torch.from_numpy(np.ascontiguousarray(fr[k])).pin_memory() for k in FRAME_KEYS}
    resident = {k: v.to(dev) for k, v in host.items()}
    img = torch.empty((H * W, 3), device=dev)
    img_host = torch.empty((H * W, 3)).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = img_host.numel() * 4
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)        # > 126 MB L2

    def step_resident():
        render_rays(net, resident, CHUNK, out=img)

    def step_e2e():
        f = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        render_rays(net, f, CHUNK, out=img)
        img_host.copy_(img, non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s, e in ev:
            flush.zero_()
            s.record()
            fn()
            e.record()
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / 1e3

    for _ in range(args.warmup):
        step_resident()
    ops.LAUNCHES = 0
    with ClockSampler(local) as clk:
        t_res = timed(step_resident, args.steps)
    launches = ops.LAUNCHES
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    t_e2e = timed(step_e2e, args.steps)
    mpix = H * W / 1e6
    value = world * args.steps * mpix / t_res
    e2e = world * args.steps * mpix / t_e2e

    # roofline of the dominant kernel: the per-neighbour MLP (4 dense layers, 542,720 FLOP per neighbour row)
    from hybridneuralrendering_b200 import profiling
    roof = profiling.dominant_kernel_roofline(net, resident, CHUNK, peaks())
    # whole-step roofline as SURVEY.md 8(d) defines it (sum over stages of max(bytes/HBM, FLOPs/tensor peak) / measured time)
    roof["step"] = profiling.step_roofline(roof["units"], V, H * W, int(opt.z_depth_dim), int(opt.SR), H, W, t_res / args.steps * 1e3, peaks()[0])

    line = {"metric": "render Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": H * W, "chunk_rays": CHUNK or H * W, "valid_samples_per_pass": net.aggregator.max_valid_chunk, "use_nearest": V, "SR": int(opt.SR), "K": int(opt.K),
                       "l2": "flushed between steps (256 MB write)", "mlp_engine": "tc (3xFP16 tcgen05 fused kernels)",
                       "parallelism": f"{world} independent frame(s), one per GPU"},
            "e2e": {"value": e2e, "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clk.summary(), "roofline": roof}
    # second half of BASELINE.json's metric: train rays/s (fwd+bwd) on configs[2]; under torchrun every rank trains on its own
    # 4096-ray batch and the gradients are all-reduced over NCCL (weak scaling), time = max over ranks
    if not args.no_train:
        from hybridneuralrendering_b200.benchmarks import train_step_benchmark
        del net, resident, flush
        torch.cuda.empty_cache()
        tr = train_step_benchmark(dev, steps=max(3, args.steps), warmup=max(3, args.warmup), world=world, rank=rank, stage_split=(world == 1))
        tt = torch.tensor([tr["ms_fwd_bwd"]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tr["ms_fwd_bwd"] = float(tt.item())
        tr["value"] = world * tr["rays"] / (tr["ms_fwd_bwd"] * 1e-3)
        tr["n_gpus"] = world
        tr["roofline"] = profiling.step_roofline(tr, tr["views"], tr["rays"], 400, 24, 480, 640, tr["ms_fwd_bwd"], peaks()[0], train=True,
                                                 points=tr["points"])
        tr["gradient_allreduce"] = "NCCL SUM, dense point gradients + coalesced MLP bucket" if world > 1 else "none (1 GPU)"
        line["train"] = tr
        if world == 1:
            from hybridneuralrendering_b200.benchmarks import blur_train_step_benchmark
            torch.cuda.empty_cache()
            line["train_blur"] = blur_train_step_benchmark(dev, steps=3, warmup=3)      # BASELINE configs[3]
            line["train_blur_learnable"] = blur_train_step_benchmark(dev, steps=3, warmup=3, learnable=True)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        q_np, n_sample = sample_query(xyz, att, fr, P, dev)
        cores = os.cpu_count() or 1
        _, dt = cpu_reference_sample(P, xyz, att, fr, opt, q_np, cores)
        line["cpu_baseline"] = {"value": n_sample / dt / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "port",
                                "sample": f"{n_sample} kept rays of the same frame, aggregation+compositing (reference's torch CPU path restated), {dt:.1f} s"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
