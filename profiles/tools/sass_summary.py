"""Per-kernel counts of the Blackwell-specific SASS instructions in the built library.

    python profiles/tools/sass_summary.py > profiles/r02_sass_summary.txt

cuobjdump -sass on libexaspim_b200.so: UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM (tcgen05.ld,
TMEM -> registers), UTMALDG / UTMASTG (TMA tensor load / store), UTCBAR (tcgen05.commit), SYNCS
(mbarrier), HMMA (legacy mma.sync -- must be 0 on the hot path), FFMA / HFMA2 for the SIMT kernels."""

import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "aind_exaspim_neuron_segmentation_b200", "libexaspim_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "UTCBAR", "SYNCS", "HMMA",
       "FFMA", "STG", "LDG", "ATOMG", "REDG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {op: 0 for op in OPS}
            kernels[cur]["instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["instructions"] += 1
        base = op.split(".")[0]
        if base in kernels[cur]:
            kernels[cur][base] += 1
        if base in ("UTCHMMA", "UTMALDG") and ".2CTA" in op:
            kernels[cur][base + ".2CTA"] += 1
    names = list(kernels)
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, out):
        demangle[n] = d
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a); counts are static SASS instructions")
    print("# " + " | ".join(["kernel", "instructions"] + OPS))
    tot = {op: 0 for op in OPS}
    for n in sorted(names, key=lambda k: demangle[k]):
        k = kernels[n]
        short = re.sub(r"\(.*\)$", "", demangle[n]).replace("void ", "").replace("exa::", "")
        print(" | ".join([short, str(k["instructions"])] + [str(k[op]) for op in OPS]))
        for op in OPS:
            tot[op] += k[op]
    print(" | ".join(["TOTAL", "-"] + [str(tot[op]) for op in OPS]))


if __name__ == "__main__":
    sys.exit(main())
