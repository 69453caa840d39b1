"""SASS opcode evidence per kernel of the shipped library (no GPU needed):
    python scripts/sass_histogram.py > profiles/r02_sass_histogram.json
Counts the mnemonics that prove the Blackwell paths (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG /
UTMAREDG = TMA load / store / reduce, FFMA2 / FADD2 / FMUL2 = packed fp32) next to the totals."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "quantool_b200", "lib", "libquantool_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "FFMA2", "FADD2", "FMUL2",
        "FFMA", "FADD", "FMUL", "FMNMX", "MUFU", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "SHFL", "LDGSTS", "BAR"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("qt::", "")
            cur = out.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["total"] += 1
            op = m.group(1)
            for k in KEYS:
                if op == k or (k in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM") and op.startswith(k)):
                    cur[k] += 1
    res = {k: {kk: v[kk] for kk in ["total"] + KEYS if v[kk]} for k, v in sorted(out.items())}
    summary = {"library": os.path.relpath(SO, ROOT), "kernels": len(res),
               "tcgen05_kernels": sorted(k for k, v in res.items() if v.get("UTCHMMA")),
               "tma_kernels": sorted(k for k, v in res.items() if v.get("UTMALDG") or v.get("UTMASTG") or v.get("UTMAREDG"))}
    json.dump({"summary": summary, "per_kernel": res}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
