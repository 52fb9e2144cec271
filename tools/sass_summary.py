"""Per-kernel SASS evidence of libsfb200.so: registers / shared memory (cuobjdump -res-usage) and the instruction counts
that prove the hardware paths (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store,
UTCBAR = tcgen05.commit, MUFU.EX2, SYNCS = mbarrier ops, STL / LDL = register spills).  No GPU needed.

    python tools/sass_summary.py > profiles/r02_sass_summary.json
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "self_forcing_b200", "libsfb200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "MUFU.EX2", "MUFU.TANH", "SYNCS",
             "STL", "LDL", "USETMAXREG", "HMMA", "FFMA2", "ELECT"]


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return r.stdout.splitlines()


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = dict(registers=int(m.group(2)), stack=int(m.group(3)), static_smem=int(m.group(4)), local=int(m.group(5)))
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        body = line.split("*/", 1)[-1]
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["instructions"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                kernels[cur][mn] += 1
    names = list(kernels)
    pretty = dict(zip(names, demangle(names)))
    out = {"_how": "cuobjdump -sass / -res-usage of self_forcing_b200/libsfb200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a); "
                   "counts are static instruction counts per kernel; tools/sass_summary.py", "kernels": {}}
    tot = collections.Counter()
    for k, cnt in kernels.items():
        name = re.sub(r"\(.*", "", pretty[k]).replace("void ", "")
        rec = dict(cnt)
        rec.update(usage.get(k, {}))
        out["kernels"][name] = rec
        tot.update({m: cnt[m] for m in MNEMONICS})
    out["totals"] = {m: tot[m] for m in MNEMONICS if tot[m]}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
