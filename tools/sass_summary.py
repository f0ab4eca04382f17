"""profiles/r2_sass_summary.txt: per kernel of libdlc.so, the count of the SASS mnemonics that prove the Blackwell
paths (tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA loads / stores -> UTMALDG / UTMASTG) - from
`cuobjdump -sass deeploopcloser_b200/libdlc.so` (no GPU needed).    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "deeploopcloser_b200", "libdlc.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
PAT = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"),
                               ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.2CTA", r"\bUTMALDG[.\w]*\.2CTA"),
                               ("UTMALDG.IM2COL", r"\bUTMALDG[.\w]*IM2COL"), ("UTMALDG.3D", r"\bUTMALDG\.3D"),
                               ("UTMASTG", r"\bUTMASTG"), ("HMMA (legacy)", r"\bHMMA")])
kern, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("dlc::", "").replace("void ", "")
        counts[kern] = collections.Counter()
        continue
    if kern:
        for name, pat in PAT.items():
            if re.search(pat, line):
                counts[kern][name] += 1
print("SASS mnemonic counts per kernel of deeploopcloser_b200/libdlc.so (cuobjdump -sass, sm_100a)")
print("%-64s " % "kernel" + " ".join("%14s" % n for n in PAT))
tot = collections.Counter()
for k, c in counts.items():
    if not any(c.values()):
        continue
    print("%-64s " % k[:64] + " ".join("%14d" % c[n] for n in PAT))
    tot.update(c)
print("%-64s " % "TOTAL" + " ".join("%14d" % tot[n] for n in PAT))
print("kernels in the library: %d, of which %d use tcgen05 / TMA" % (len(counts), sum(1 for c in counts.values() if any(c.values()))))
