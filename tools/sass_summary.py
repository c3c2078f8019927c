#!/usr/bin/env python
"""Per-kernel SASS evidence for libnkb200.so: counts of the FP64 tensor instruction (DMMA.8x8x4), of the Blackwell bulk-copy
engine instructions (UBLKCP = cp.async.bulk, UBLKRED = cp.reduce.async.bulk, UTMALDG = tensor-map TMA), of Ampere-style
LDGSTS (cp.async), and of tcgen05 / TMEM instructions (UTC*MMA, LDTM -- none expected: Blackwell has no FP64 tcgen05 path).

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
LIB = ROOT / "nys_koop_lqr_b200" / "libnkb200.so"
PATTERNS = collections.OrderedDict([
    ("DMMA.8x8x4", re.compile(r"\bDMMA\.8x8x4")), ("UBLKCP", re.compile(r"\bUBLKCP")), ("UBLKRED", re.compile(r"\bUBLKRED")),
    ("UTMALDG", re.compile(r"\bUTMALDG")), ("LDGSTS", re.compile(r"\bLDGSTS")), ("SYNCS (mbarrier)", re.compile(r"\bSYNCS")),
    ("UTC*MMA / LDTM", re.compile(r"\b(UTC\w*MMA|LDTM|STTM)")), ("BAR.SYNC", re.compile(r"\bBAR\.SYNC")),
    ("DFMA", re.compile(r"\bDFMA")), ("MUFU", re.compile(r"\bMUFU")),
])


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        mt = re.match(r"\s*Function : (\S+)", line)
        if mt:
            cur = mt.group(1)
            kernels[cur] = collections.Counter()
            kernels[cur]["_lines"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", line):
            kernels[cur]["_lines"] += 1
            for name, pat in PATTERNS.items():
                if pat.search(line):
                    kernels[cur][name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS summary of `nys_koop_lqr_b200/libnkb200.so` (sm_100a) -- regenerate with `python tools/sass_summary.py`\n")
    print("Counts are static instruction counts per kernel from `cuobjdump -sass`.  `DMMA.8x8x4` is the FP64 tensor-core instruction")
    print("(Blackwell has no FP64 tcgen05/TMEM path, so UTC*MMA / LDTM are expected to be 0 everywhere); `UBLKCP` / `UBLKRED` are the")
    print("bulk-copy (TMA) engine's global->shared copy and shared->global reduce-add; `LDGSTS` is the per-thread cp.async.\n")
    cols = list(PATTERNS)
    print("| kernel | SASS lines | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for (mangled, cnt), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name)
        print(f"| `{short}` | {cnt['_lines']} | " + " | ".join(str(cnt[c]) for c in cols) + " |")
    return 0


if __name__ == "__main__":
    sys.exit(main())
