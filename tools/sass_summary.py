"""profiles/r02_sass_summary.txt: per kernel of the built library, the SASS mnemonics that prove the Blackwell-native
paths (tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UBLKCP / UTMALDG / UTMASTG), instruction count, registers.
usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt   (needs only cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "emap_b200", "lib", "libemap_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+)", ln)
    if m and cur:
        regs[cur] = int(m.group(1))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
counts, order, cur = {}, [], None
for ln in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        op = m.group(1)
        counts[cur]["instructions"] += 1
        for key in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "MUFU", "SYNCS"):
            if op.startswith(key):
                counts[cur][key] += 1
print("kernel | instr | regs | UTCHMMA (tcgen05.mma) | LDTM (tcgen05.ld) | UBLKCP (1-D bulk copy) | UTMALDG (TMA tensor load) | UTMASTG (TMA tensor store) | "
      "UTCBAR (tcgen05.commit) | SYNCS (mbarrier) | MUFU | HMMA (legacy mma.sync)")
for fn in sorted(order, key=lambda f: -counts[f]["instructions"]):
    c = counts[fn]
    name = demangle(fn)
    name = re.sub(r"\(.*$", "", name)[:90]
    print(f"{name} | {c['instructions']} | {regs.get(fn, '?')} | {c['UTCHMMA']} | {c['LDTM']} | {c['UBLKCP']} | "
          f"{c['UTMALDG']} | {c['UTMASTG']} | {c['UTCBAR']} | {c['SYNCS']} | {c['MUFU']} | {c['HMMA']}")
