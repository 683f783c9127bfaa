"""Static counts of the Blackwell-specific SASS mnemonics per kernel of libcoverb200.so (cuobjdump -sass | c++filt).
usage: python tools/sass_mnemonics.py > profiles/<tag>_sass_mnemonics.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

SO = Path(__file__).resolve().parent.parent / "cover_vla_b200" / "libcoverb200.so"
WANT = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UCGABAR_ARV", "UCGABAR_WAIT", "LDGSTS",
        "HMMA", "SYNCS", "ACQBULK", "UTCATOMSWS")

sass = subprocess.run(f"cuobjdump -sass {SO} | c++filt", shell=True, capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (.*)$", line)
    if m:
        cur = re.sub(r"\(.*$", "", m.group(1).replace("(anonymous namespace)", "<unnamed>")).strip()
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(?:\.|\s|;)", line)
    if m and m.group(1) in WANT:
        counts[cur][m.group(1)] += 1
print("# cuobjdump -sass cover_vla_b200/libcoverb200.so: Blackwell-specific SASS mnemonics per kernel (static instruction counts)")
print("# UTCHMMA = tcgen05.mma (cta_group::1 and ::2), LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = bulk copy,")
print("# UCGABAR = cluster barrier, SYNCS = mbarrier ops, LDGSTS = cp.async, HMMA = legacy mma.sync (fallback attention kernels only)")
tot = collections.Counter()
for k, c in counts.items():
    if not c:
        continue
    tot.update(c)
    print(f"{k[:92]:92s} " + " ".join(f"{n}={c[n]}" for n in sorted(c)))
print("# total: " + " ".join(f"{n}={tot[n]}" for n in sorted(tot)))
