"""Summarise `ncu -i X.ncu-rep --page source --csv` (stdin): per kernel, the warp-stall totals, the
totals per SASS opcode and the hottest instructions with two instructions of context (the .ncu-rep
files are too big to bring back from the GPU box, this text summary is what travels)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}
        secs.append(cur)
    elif cur is not None:
        if cur["hdr"] is None:
            cur["hdr"] = r
        else:
            cur["rows"].append(r)
seen = set()
for sec in secs:
    if sec["name"] in seen:
        continue
    seen.add(sec["name"])
    hdr = sec["hdr"]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    byop = collections.defaultdict(collections.Counter)
    inst = []
    n = 0
    for k, r in enumerate(sec["rows"]):
        if len(r) < len(hdr):
            continue
        try:
            ns = int(r[idx["# Samples"]])
        except Exception:
            continue
        n += ns
        src = r[idx["Source"]].strip()
        op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0].split(".")[0] if src else "?"
        byop[op]["_n"] += ns
        c = collections.Counter()
        for s_ in stalls:
            v = int(r[idx[s_]] or 0)
            if v:
                tot[s_] += v
                byop[op][s_] += v
                c[s_] += v
        inst.append((ns, k, src, c))
    print("=====", sec["name"], "samples", n)
    for k, v in tot.most_common(8):
        print(f"   {k:26s} {v:9d} {100 * v / max(n, 1):5.1f}%")
    for op, c in sorted(byop.items(), key=lambda kv: -kv[1]["_n"])[:14]:
        m = c.pop("_n")
        print(f"   op {op:10s} {m:9d} {100 * m / max(n, 1):5.1f}%  {dict(c.most_common(3))}")
    print("   -- hottest instructions (samples, share, SASS; top stalls) with 2 preceding instructions")
    srcs = {k: s for _, k, s, _ in inst}
    for ns, k, src, c in sorted(inst, key=lambda t: -t[0])[:28]:
        ctx = " <- ".join(srcs.get(k - d, "")[:44] for d in (1, 2))
        print(f"   {ns:8d} {100 * ns / max(n, 1):5.1f}%  {src[:70]:70s} {dict(c.most_common(2))}  [{ctx}]")
