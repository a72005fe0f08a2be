#!/usr/bin/env python
"""Turn an ncu report (or a launch-list csv) into the markdown summaries committed under profiles/.

    python scripts/summarize_ncu.py rep   gpurun_out/prof_x.ncu-rep  "title" "command"  > profiles/x.md
    python scripts/summarize_ncu.py list  gpurun_out/launches.csv    "title" "command" [frames] > profiles/y.md
    python scripts/summarize_ncu.py metrics gpurun_out/membound.csv  "title" "command" > profiles/z.md   (ncu --metrics ... --csv)
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def rep(path, title, cmd):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, data = rows[0], rows[1], rows[2:]
    ki = h.index("Kernel Name")
    print(f"# {title}\n\nCommand: `{cmd}`\n(`ncu -i {path.split('/')[-1]} --page raw --csv`; one column per captured launch)\n")
    print("| metric | unit | " + " | ".join(f"launch {i + 1}" for i in range(len(data))) + " |")
    print("|---|---|" + "---|" * len(data))
    print("| kernel | | " + " | ".join(d[ki][:48] for d in data) + " |")
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"| `{k}` | {u[i]} | " + " | ".join(d[i] for d in data) + " |")


def launch_list(path, title, cmd, frames):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {title}\n\nCommand: `{cmd}`\n(cold-cache, serialised launches: compare SHARES, not absolute times; {frames} frame passes in the capture)\n")
    print("| kernel | launches | total ms | ms / frame | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        print(f"| `{k}` | {n} | {t:.2f} | {t / frames:.2f} | {100 * t / tot:.1f}% |")
    print(f"\nTotal {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches.")


def metrics(path, title, cmd, peak_gbs=None):
    """per-launch table of a `ncu --metrics a,b,c --csv` log (one csv row per launch and metric)"""
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ii, ki, mi, vi, ui = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    launches = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        L = launches.setdefault(r[ii], {"kernel": re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "").replace("void ", ""),
                                        "grid": r[gi], "block": r[bi]})
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        if r[mi].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
        L[r[mi]] = v
    print(f"# {title}\n\nCommand: `{cmd}`\n(per-launch counters, cold caches, serialised; `dram GB/s` = (read + written bytes) / duration)\n")
    print("| # | kernel | grid x block | regs | us | DRAM read MB | DRAM written MB | dram GB/s | dram % of peak | L2 hit % | SM thr % | warps active % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for i, L in launches.items():
        us = L.get("gpu__time_duration.sum", 0.0)
        rd, wr = L.get("dram__bytes_read.sum", 0.0), L.get("dram__bytes_write.sum", 0.0)
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us else 0.0
        print(f"| {i} | `{L['kernel'][:60]}` | {L['grid'].split(',')[0][1:]} x {L['block'].split(',')[0][1:]} | "
              f"{L.get('launch__registers_per_thread', 0):.0f} | {us:.1f} | {rd / 1e6:.2f} | {wr / 1e6:.2f} | {gbs:.0f} | "
              f"{L.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):.1f} | {L.get('lts__t_sector_hit_rate.pct', 0):.1f} | "
              f"{L.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):.1f} | "
              f"{L.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.1f} |")


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        rep(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "metrics":
        metrics(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        launch_list(sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5]) if len(sys.argv) > 5 else 1.0)
