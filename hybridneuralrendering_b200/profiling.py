"""Live device timing of the hot-path kernels (CUDA events on the launching stream) and the roofline
record bench.py publishes.  Algorithmic work per unit: SURVEY.md §8(d) / DESIGN.md."""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from .renderer import render_rays

FLOP_PER_NEIGHBOUR = 542_720            # 2*(284*256 + 256*256 + 263*256 + 256*256 + 256), SURVEY.md §8d
DRAM_BYTES_PER_LAUNCH_NCU = 337_587_200  # measured, see profiles/r2_nbr_mlp_f16_ncu.md (90.4 MB read + 247.2 MB written)


def stage_times(net, frame, chunk_rays: int) -> Dict[str, float]:
    """render one frame with per-launch events; returns {tag: milliseconds} plus '_units' counters"""
    torch.cuda.synchronize()
    ops.TIMERS = []
    units = {"valid_neighbours": 0, "valid_samples": 0, "kept_rays": 0}
    from .renderer import render_rays
    try:
        orig = net.forward

        def counted(*a, **k):
            o = orig(*a, **k)
            ex = net.last_extras
            units["valid_samples"] += ex.n_valid
            units["kept_rays"] += ex.n_rays
            units["valid_neighbours"] += net.aggregator.last_valid_neighbours() if ex.n_valid else 0
            return o

        net.forward = counted
        try:
            render_rays(net, frame, chunk_rays)
        finally:
            net.forward = orig
        torch.cuda.synchronize()
        out: Dict[str, float] = {}
        launches: Dict[str, int] = {}
        for tag, s, e in ops.TIMERS:
            out[tag] = out.get(tag, 0.0) + s.elapsed_time(e)
            launches[tag] = launches.get(tag, 0) + 1
    finally:
        ops.TIMERS = None
    out["_units"] = units
    out["_launches"] = launches
    return out


def dominant_kernel_roofline(net, frame, chunk_rays: int, peaks_and_kind) -> Dict:
    """roofline record for the dominant kernel of the render step: the per-neighbour MLP.
    achieved = 542,720 FLOP x valid neighbours of the frame / summed duration of its launches."""
    peaks, kind = peaks_and_kind
    t = stage_times(net, frame, chunk_rays)
    units = t.pop("_units")
    launches = t.pop("_launches")
    total = sum(t.values())
    ms = t.get("nbr_mlp", 0.0)
    flops = FLOP_PER_NEIGHBOUR * units["valid_neighbours"]
    achieved = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    # fp32-accurate tensor-core path = 3 FP16 MMAs per product (csrc/nbr_mlp_f16.cu) -> peak = dense fp16/bf16 tensor
    # throughput / 3.  MEASURED_PEAKS.json: cuBLAS bf16 (same tcgen05 kind::f16 rate); the kernel is timed inside a long
    # step, so the sustained figure applies.
    engine = getattr(net.aggregator, "mlp_engine", "tc")
    if engine == "tc":
        peak = peaks["bf16_tflops_sustained"] / 3.0
        basis = f"bf16_tflops_sustained/3 of {kind} (3xFP16 split, fp32-equivalent FLOPs)"
    else:
        peak = peaks["bf16_tflops_sustained"] / 2.0 / 3.0
        basis = f"bf16_tflops_sustained/2/3 of {kind} (3xTF32, TF32 dense = bf16/2 derived)"
    nl = launches.get("nbr_mlp", 0)
    return {"bound": "tensor", "kernel": "nbr_mlp_f16_kernel: fused per-neighbour MLP (gather + block1 + block3 + density head + K-sum)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            "traffic": DRAM_BYTES_PER_LAUNCH_NCU, "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of 262,144 valid "
            "samples, ncu --set full (profiles/r2_nbr_mlp_f16_ncu.md); algorithmic HBM bytes per launch = 168 B x neighbours in + 1124 B x samples out",
            "algorithmic_flop_per_launch": FLOP_PER_NEIGHBOUR * units["valid_neighbours"] / nl if nl else None,
            "peak_basis": basis, "launches": nl, "kernel_ms_per_step": ms, "kernel_ms_per_launch": ms / nl if nl else None,
            "share_of_step": ms / total if total else None,
            "units": units, "stage_ms": {k: round(v, 3) for k, v in sorted(t.items())}}


def step_roofline(units: Dict[str, int], views: int, rays: int, depth_samples: int, SR: int, height: int, width: int, measured_ms: float,
                  peaks: Dict[str, float], train: bool = False, points: int = 0) -> Dict:
    """Roofline time of a whole step as SURVEY.md 8(d) defines it: sum over the stages of max(algorithmic bytes / HBM peak,
    algorithmic FLOPs / tensor peak); `frac` = that / the measured CUDA-event time of the step.

    Bytes and FLOPs per unit are the table of SURVEY.md 8(d) (compulsory traffic only).  Counts: M = valid neighbour rows,
    Nv = valid samples, R'' = kept rays (all measured on this run's inputs).  The candidate-point term `16*Cand` of the neighbour
    search is not counted by the product and is left out, which only LOWERS the roofline time (conservative fraction).
    Tensor peaks: the fp32-accurate forward runs 3 FP16 MMAs per product (peak = bf16_tflops_sustained / 3); since round 2 the fused
    backward runs on the same kind::f16 pipe (3 BF16 MMAs per product), so its peak is the same -- `frac`.  `frac_tf32_basis` keeps
    round 1's denominator for the backward (3 TF32 MMAs per product; TF32 dense sustained as MEASURED by scripts/measure_tf32_peak.py
    when profiles/r2_tf32_peak.json exists, else bf16 / 2) so that the two rounds stay comparable.  train=True adds the backward rows (FLOPs x2, gather recompute + scatter-add RMW,
    image-gather scatter, compositing backward, pyramid dgrad/wgrad) and the dense zero-filled point-gradient tables (39*N floats)."""
    M, Nv, Rk = int(units["valid_neighbours"]), int(units["valid_samples"]), int(units["kept_rays"])
    V = int(views)
    hbm = float(peaks["hbm_gbs"]) * 1e9
    t_fwd = float(peaks["bf16_tflops_sustained"]) / 3.0 * 1e12
    t_bwd = t_fwd
    t_bwd_tf32 = float(peaks.get("tf32_tflops_sustained", float(peaks["bf16_tflops_sustained"]) / 2.0)) / 3.0 * 1e12
    nbr_flop = FLOP_PER_NEIGHBOUR * M
    smp_flop = (154_184 + 39_040 * V) * Nv
    st = {   # stage: (bytes, fwd flops, bwd flops)
        "query (ray generation, occupancy mask, sample selection, neighbour search)":
            (24 * rays + 4 * rays * depth_samples + 16 * Rk * SR + 124 * Nv, 0, 0),
        "gather + weights": (168 * M + 24 * Nv + 16 * Nv + ((16 * Nv + 168 * M + 2 * 156 * M) if train else 0), 0, 0),
        "per-neighbour MLP": (0, nbr_flop, 2 * nbr_flop if train else 0),
        "per-sample MLPs": (0, smp_flop, 2 * smp_flop if train else 0),
        "image gather": (200 * V * Nv + (2 * 180 * V * Nv if train else 0), 0, 0),
        "feature pyramid (cuDNN)": ((12 + 10.5) * V * height * width * (3 if train else 1), 0, 0),
        "compositing": ((21 + 8) * Rk * SR + 16 * Rk + (((21 + 8) * Rk * SR + 32 * Rk + 16 * Rk * SR) if train else 0), 0, 0),
    }
    if train and points:
        st["dense point-gradient tables (zero fill)"] = (39 * 4 * points, 0, 0)
    rows, total, total_tf32 = {}, 0.0, 0.0
    for k, (b, f, g) in st.items():
        ms = max(b / hbm, f / t_fwd + g / t_bwd) * 1e3
        rows[k] = {"bytes": int(b), "flop": int(f + g), "roofline_ms": round(ms, 4), "bound": "tensor" if f else "hbm"}
        total += ms
        total_tf32 += max(b / hbm, f / t_fwd + g / t_bwd_tf32) * 1e3
    return {"definition": "sum over stages of max(bytes / HBM peak, FLOPs / tensor peak) / measured step time (SURVEY.md 8d)",
            "roofline_ms": total, "measured_ms": measured_ms, "frac": total / measured_ms if measured_ms else None,
            "roofline_ms_tf32_basis": total_tf32, "frac_tf32_basis": total_tf32 / measured_ms if (measured_ms and train) else None,
            "hbm_peak_gbs": peaks["hbm_gbs"], "tensor_peak_fwd_tflops": t_fwd / 1e12, "tensor_peak_bwd_tflops": t_bwd / 1e12,
            "tensor_peak_bwd_tf32_basis_tflops": t_bwd_tf32 / 1e12,
            "tensor_share_of_roofline": sum(r["roofline_ms"] for r in rows.values() if r["bound"] == "tensor") / total if total else None,
            "stages": rows}
