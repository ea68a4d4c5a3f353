#!/usr/bin/env python
"""Static view of every fused-step kernel in libmpe_b200.so (no GPU needed): registers / stack from
`cuobjdump -res-usage`, instruction mix from `cuobjdump -sass`.  The counts are STATIC instructions of the whole
kernel; every entity loop is unrolled, but the runtime-selected alternatives (cp.async vs TMA action staging, the
partial-warp tail path with scalar loads / stores, exact-image vs padded tile streaming) are all in the binary, so a
full warp executes roughly half of them (spread N=3: 664 executed per warp in ncu vs 1252 static).

    python tools/sass_stats.py > profiles/r1_static_resources.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("MPE_B200_LIB", os.path.join(ROOT, "multiagent_particle_envs_b200", "csrc", "libmpe_b200.so"))


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", ln)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    mix = {}
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            mix[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", ln)
        if m and cur:
            op = m.group(1)
            if op == "NOP":
                continue
            mix[cur]["total"] += 1
            mix[cur][op] += 1
            if op == "MUFU":
                mix[cur]["MUFU" + (m.group(2) or "")] += 1
    names = demangle(sorted(usage))
    rows = []
    for mangled, (reg, stack) in usage.items():
        nm = names[mangled]
        m = re.match(r"void mpe::mpe_kernel<mpe::(.+), 0, (false|true), (false|true), (false|true)>\(", nm)
        if not m:
            continue
        label = (m.group(1) + (" (warp pair)" if m.group(2) == "true" else "") + (" HOT" if m.group(3) == "true" else "")
                 + (" 80-reg" if m.group(4) == "true" else ""))
        c = mix.get(mangled, {})
        fp = sum(c[k] for k in ("FADD", "FMUL", "FFMA", "FSETP", "FSEL", "FMNMX", "FMNMX3"))
        rows.append((label, reg, stack, c["total"], fp, c["MUFU"], c["LDG"], c["LDGSTS"], c["LDS"], c["STS"], c["STG"],
                     c["BRA"] + c["BSSY"] + c["BSYNC"], c["WARPSYNC"] + c["BAR"]))
    rows.sort(key=lambda r: r[3])
    print("# Static resources of the fused-step kernels (`tools/sass_stats.py`, sm_100a, %s)\n" % os.path.basename(LIB))
    print("`__launch_bounds__(512, 1)` caps registers at 128.  No kernel spills (STACK 0).  Columns are static SASS counts of "
          "the whole kernel including the runtime-selected alternatives (TMA vs cp.async staging, partial-warp tail), NOPs excluded; `fp` = FADD+FMUL+FFMA+FSETP+FSEL+FMNMX, `branch` = BRA+BSSY+BSYNC.\n")
    print("| program | regs | stack | instr | fp | MUFU | LDG | LDGSTS | LDS | STS | STG | branch | sync |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print("| " + " | ".join(str(x) for x in r) + " |")
    other = [(names[k], v) for k, v in usage.items() if "mpe_kernel" not in names[k] or ", 0, " not in names[k]]
    spills = [n for n, (r, s) in other if s]
    print("\nOther kernels with a non-zero stack frame: %s" % (", ".join("`%s`" % s for s in spills) or "none"))


if __name__ == "__main__":
    sys.exit(main())
