#!/usr/bin/env python
"""Per-kernel count of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
UTMALDG = TMA loads, LDTM / STTM = tcgen05.ld / st, UBLKPF = bulk L2 prefetch, plus HMMA (legacy mma.sync: must be 0).

  python tools/sass_summary.py > profiles/r02_sass_summary.txt      (cuobjdump -sass on the built libpg_b200.so)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200", "pg_b200", "libpg_b200.so")
PAT = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UBLKPF", "UBLKCP", "HMMA", "LDGSTS",
       "STG.E.64.STRONG.SYS", "LDG.E.128.STRONG.SYS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            per.setdefault(name, collections.Counter())
            continue
        if name is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            c = per[name]
            c["instructions"] += 1
            if op.startswith("UTCHMMA"):
                c["UTCHMMA"] += 1
                if ".2CTA" in op:
                    c["UTCHMMA.2CTA"] += 1
            for p in ("UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UBLKPF", "UBLKCP", "HMMA", "LDGSTS"):
                if op.startswith(p):
                    c[p] += 1
            if op.startswith("STG.E.64.STRONG.SYS"):
                c["STG.E.64.STRONG.SYS"] += 1
            if op.startswith("LDG.E.128.STRONG.SYS"):
                c["LDG.E.128.STRONG.SYS"] += 1
    fam = collections.OrderedDict()
    for k, c in per.items():
        base = re.sub(r"<.*", "", k).split("::")[-1]
        f = fam.setdefault(base, collections.Counter())
        f.update(c)
        f["variants"] += 1
    cols = ["variants", "instructions"] + PAT
    print("kernel".ljust(36) + " ".join(c.rjust(12)[:12] for c in cols))
    tot = collections.Counter()
    for k, c in sorted(fam.items()):
        print(k[:35].ljust(36) + " ".join(str(c.get(col, 0)).rjust(12) for col in cols))
        tot.update(c)
    print("TOTAL".ljust(36) + " ".join(str(tot.get(col, 0)).rjust(12) for col in cols))
    print("\nSTG.E.64.STRONG.SYS / LDG.E.128.STRONG.SYS: the {value, sequence} stores into peer exchange buffers and the volatile "
          "16-byte polls of the tensor-parallel exchange (csrc/tp_exchange.cuh) inside the GEMV / norm kernels.")


if __name__ == "__main__":
    main()
