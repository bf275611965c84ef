#!/usr/bin/env python
"""Per-kernel SASS opcode evidence for libsdxl_b200.so (no GPU needed): `python tools/sass_histogram.py > profiles/<tag>_sass_histogram.txt`.

For every kernel in the built library: instruction count and the counts of the mnemonics that prove the Blackwell-native path
(B200_PROFILING.md "What proves a Blackwell-native kernel"): UTC*MMA (tcgen05.mma; .2CTA = cta_group::2), LDTM / STTM
(tcgen05.ld / .st), UTMALDG / UTMASTG (TMA tensor loads / stores; .MULTICAST), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
MUFU / F2FP (the XU pipe the attention softmax is bound by), HMMA (legacy mma.sync: must be 0), LDG / STG widths."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stable-diffusion-xl-burn_b200", "sdxl_b200", "libsdxl_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCHMMA tmem-A", "LDTM", "STTM", "UTMALDG", "UTMALDG.MULTICAST", "UTMASTG", "UTCBAR", "SYNCS", "ELECT",
        "R2UR", "MUFU.EX2", "F2FP", "HMMA", "LDG.E.128", "STG.E.128", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*);", line)
        if not (cur and m):
            continue
        op, rest = m.group(1), m.group(2)
        c = kernels[cur]
        c["instructions"] += 1
        base = op.split(".")[0]
        if base == "UTCHMMA":
            c["UTCHMMA"] += 1
            if ".2CTA" in op:
                c["UTCHMMA.2CTA"] += 1
            if rest.strip().startswith("tmem["):
                c["UTCHMMA tmem-A"] += 1
        elif base in ("LDTM", "STTM", "UTMASTG", "UTCBAR", "SYNCS", "ELECT", "R2UR", "F2FP", "HMMA", "BAR"):
            c[base] += 1
        elif base == "UTMALDG":
            c["UTMALDG"] += 1
            if "MULTICAST" in op:
                c["UTMALDG.MULTICAST"] += 1
        elif op.startswith("MUFU.EX2"):
            c["MUFU.EX2"] += 1
        elif op.startswith("LDG.E.128") or op.startswith("STG.E.128"):
            c[op[:9]] += 1
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); columns: {', '.join(KEYS)}")
    print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{k[:9]:>9s}" for k in KEYS))
    for name, c in kernels.items():
        print(f"{name[:58]:58s} {c['instructions']:6d} " + " ".join(f"{c[k]:9d}" for k in KEYS))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print(f"{'TOTAL':58s} {tot['instructions']:6d} " + " ".join(f"{tot[k]:9d}" for k in KEYS))
    assert tot["HMMA"] == 0, "legacy mma.sync found"


if __name__ == "__main__":
    main()
