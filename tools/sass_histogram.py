"""SASS evidence for the shipped library: per-kernel opcode histogram from `cuobjdump -sass` of libjpgenc_b200.so.
    python tools/sass_histogram.py [out.json]
Writes the counts of the opcodes the design rests on (bulk/TMA copies, packed FP32x2, warp match/vote, FP64) plus the
ten most frequent opcodes of every kernel."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jpgenc_b200", "lib", "libjpgenc_b200.so")
KEY = ["UBLKCP", "UBLKPF", "UTMASTG", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "PRMT", "VIMNMX", "MATCH",
       "VOTE", "REDUX", "SHFL", "ATOMS", "ATOMG", "RED", "DADD", "DMUL", "DFMA", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "POPC", "FLO", "LOP3", "SHF"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            base, k = name, 2
            while name in kernels:                    # template instances demangle to the same prefix
                name = f"{base}#{k}"; k += 1
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    res = {}
    for name, c in kernels.items():
        tot = sum(c.values())
        res[name] = {"instructions": tot, "key_opcodes": {k: c[k] for k in KEY if c.get(k)}, "top10": dict(c.most_common(10))}
    doc = {"source": "cuobjdump -sass jpgenc_b200/lib/libjpgenc_b200.so (sm_100a), static instruction counts per kernel",
           "nvcc": subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2],
           "kernels": res}
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_opcodes.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    for name, e in res.items():
        print(f"{name}: {e['instructions']} instr; " + ", ".join(f"{k} {v}" for k, v in e["key_opcodes"].items() if k in KEY[:17]))


if __name__ == "__main__":
    main()
