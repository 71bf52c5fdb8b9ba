#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-only SASS opcodes in libtcn_b200.so (cuobjdump -sass):
UTCHMMA / UTCQMMA / UTCIMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCCP (tcgen05.cp), UTMALDG / UTMASTG (TMA),
UTCBAR (tcgen05.commit), HMMA / IMMA (legacy mma.sync), SYNCS (mbarrier).  Writes profiles/r2_sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "computervision_codes_b200", "csrc", "libtcn_b200.so")
OPS = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "SYNCS")


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for op in OPS:
            if re.search(r"\b" + op + r"[\.\s]", line):
                counts[cur][op] += 1
    dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    names = {k: (re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", d)[:100] if d else k) for k, d in zip(counts, dem)}
    lines = ["# cuobjdump -sass computervision_codes_b200/csrc/libtcn_b200.so (sm_100a): Blackwell-only opcodes per kernel",
             "# " + " ".join(f"{op:>8s}" for op in OPS) + "  kernel"]
    tot = collections.Counter()
    for k, c in counts.items():
        tot.update(c)
        lines.append("  " + " ".join(f"{c[op]:8d}" for op in OPS) + "  " + names[k])
    lines.append("  " + " ".join(f"{tot[op]:8d}" for op in OPS) + "  TOTAL")
    txt = "\n".join(lines) + "\n"
    dst = os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
    open(dst, "w").write(txt)
    sys.stdout.write(txt)


if __name__ == "__main__":
    main()
