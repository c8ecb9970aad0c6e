"""Per-kernel SASS evidence that the conv path is tcgen05 + TMEM + TMA (B200_PROFILING.md mnemonics):
  python tools/sass_summary.py [profiles/r02_sass_summary.txt]
counts UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), UTMAPF (TMA prefetch), LDTM (tcgen05.ld from TMEM), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), HMMA / IMMA (legacy mma.sync: expected 0) per kernel of kfnet_b200/libkfnet_b200.so."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "kfnet_b200", "libkfnet_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTMALDG", "UTMAPF", "LDTM", "UTCBAR", "SYNCS", "HMMA", "IMMA", "FFMA", "ATOMG", "RED"]
counts, cur, arch = collections.OrderedDict(), None, set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur:
        for mn in MN:
            if re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", line):
                counts[cur][mn] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
lines = ["SASS summary of kfnet_b200/libkfnet_b200.so (cuobjdump -sass), arch = %s" % ", ".join(sorted(arch)),
         "%-78s " % "kernel" + " ".join("%8s" % m for m in MN)]
tot = collections.Counter()
for (k, c), name in zip(counts.items(), demangled):
    name = name.replace("(anonymous namespace)::", "").replace("kfn::", "")
    name = re.sub(r"\(.*", "", name)
    lines.append("%-78s " % name[:78] + " ".join("%8d" % c[m] for m in MN))
    tot.update(c)
lines.append("%-78s " % "TOTAL" + " ".join("%8d" % tot[m] for m in MN))
text = "\n".join(lines) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text)
