"""Per-kernel SASS evidence of the Blackwell-native paths (runs WITHOUT a GPU: `cuobjdump` on the built library).

    python scripts/sass_summary.py > profiles/r2_sass_summary.md

For every kernel of libhnr.so: instruction count, registers per thread, static shared memory, and the number of tcgen05 / TMEM /
bulk-copy / mbarrier / reduction instructions (mnemonics as listed in /opt/skills/guides/B200_PROFILING.md): UTCHMMA = tcgen05.mma
kind::f16/tf32, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTCATOMSWS = TMEM alloc / dealloc, UBLKCP = cp.async.bulk, SYNCS =
mbarrier ops, REDG = red.global, LDGMC = multimem.ld_reduce, ELECT = elect.sync.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hybridneuralrendering_b200", "libhnr.so")
MN = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "REDG", "LDGMC", "ELECT", "FFMA", "MUFU"]


def run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True, check=True).stdout


def main():
    sass = run("cuobjdump", "-sass", LIB)
    res = run("cuobjdump", "--dump-resource-usage", LIB)
    usage = {m.group(1).strip(): m.group(2) for m in re.finditer(r"Function ([^:]+):\n\s*(REG:\d+[^\n]*)", res)}
    rows = []
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n", 1)[0].strip()
        n = len(re.findall(r"/\*[0-9a-f]{4,6}\*/\s+[@A-Z!]", f))
        cnt = {m: len(re.findall(r"\b" + re.escape(m), f)) for m in MN}
        u = usage.get(name, "")
        reg = re.search(r"REG:(\d+)", u)
        sh = re.search(r"SHARED:(\d+)", u)
        rows.append((name, n, reg.group(1) if reg else "?", sh.group(1) if sh else "?", cnt))
    dem = run("c++filt", *[r[0] for r in rows]).splitlines()
    print("# Round 2: SASS summary of libhnr.so (sm_100a), `python scripts/sass_summary.py`\n")
    print("`cuobjdump -sass` / `--dump-resource-usage` of the library built by `__graft_entry__.build()`; no GPU involved.  Kernels that use "
          "the tensor cores first (tcgen05.mma = `UTCHMMA`, accumulators read back from TMEM with `LDTM`, operands staged by "
          "`UBLKCP` bulk copies and mbarriers `SYNCS`); `LDGMC.E.ADD.F32` = `multimem.ld_reduce` (in-switch reduction of `peer.cu`).  Dynamic shared memory "
          "(up to 223 KB for the fused kernels) is requested at launch and not part of the static figure.\n")
    print("| kernel | SASS instr. | regs | static smem B | " + " | ".join(MN) + " |")
    print("|---|---|---|---|" + "---|" * len(MN))
    key = lambda r: (-(r[4]["UTCHMMA"] > 0), -(r[4]["LDGMC"] > 0), -r[1])
    for (name, n, reg, sh, cnt), d in sorted(zip(rows, dem), key=lambda x: key(x[0])):
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        short = re.sub(r"^void ", "", d.split("(")[0])
        print(f"| `{short}` | {n} | {reg} | {sh} | " + " | ".join(str(cnt[m]) if cnt[m] else "" for m in MN) + " |")
    tot = {m: sum(r[4][m] for r in rows) for m in MN}
    print(f"\n{len(rows)} kernels; totals: " + ", ".join(f"{m} {tot[m]}" for m in MN if tot[m]))


if __name__ == "__main__":
    sys.exit(main())
