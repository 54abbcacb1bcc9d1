#!/usr/bin/env python
"""Per-kernel SASS evidence that the hot kernels of libnerfb200.so are genuine tcgen05 / TMEM / TMA code:
counts of the Blackwell mnemonics (B200_PROFILING.md) per kernel + the MMA issue loop of each tcgen05 kernel.

    python scripts/sass_excerpt.py > profiles/r02_sass_tcgen05.txt        (no GPU needed: cuobjdump reads the in-tree .so)
"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "nerf_or_nothing_b200" / "libnerfb200.so"
MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "ELECT", "FENCE.VIEW.ASYNC",
             "HMMA", "FFMA", "MUFU")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
        elif cur and re.match(r"\s*/\*[0-9a-f]{4,6}\*/", line):
            kernels[cur].append(line.rstrip())
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS of {LIB.relative_to(ROOT)} (sm_100a), cuobjdump -sass; counts of Blackwell mnemonics per kernel\n")
    print("kernel".ljust(72) + " ".join(m.rjust(8) for m in MNEMONICS[:12]) + "   instrs")
    tc = []
    for (name, body), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm.replace("nerf::(anonymous namespace)::", "").replace("nerf::", "").replace("void ", ""))
        counts = [sum(1 for l in body if re.search(r"\b" + re.escape(m) + r"\b", l)) for m in MNEMONICS]
        print(short[:71].ljust(72) + " ".join(str(c).rjust(8) for c in counts[:12]) + f"   {len(body)}")
        if counts[0]:
            tc.append((short, body))
    for short, body in tc:
        idx = [i for i, l in enumerate(body) if "UTCHMMA" in l]
        lo, hi = max(0, idx[0] - 6), min(len(body), idx[0] + 14)
        print(f"\n## {short}: first tcgen05.mma issue site ({len(idx)} UTCHMMA in the kernel)")
        for l in body[lo:hi]:
            print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
        for mn in ("UTMALDG", "UTMASTG", "LDTM", "STTM"):
            ex = next((l for l in body if mn in l), None)
            if ex:
                print(f"   ... {mn}: " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ex).strip())


if __name__ == "__main__":
    sys.exit(main())
