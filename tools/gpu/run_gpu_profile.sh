#!/bin/bash
# Bench lines + ncu evidence, exported to CSV on the box (the .ncu-rep files are too big to travel).
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train_fp32.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --mode infer --no-cpu-baseline > gpurun_out/bench_infer_fp32.json 2> gpurun_out/bench_infer.err; echo "bench infer rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --mode infer --precision fp16 --no-cpu-baseline > gpurun_out/bench_infer_fp16.json 2>/dev/null
timeout 200 python bench.py --steps 10 --warmup 3 --precision fp16 --no-cpu-baseline > gpurun_out/bench_train_fp16.json 2>/dev/null
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_kernel<3, 1|mlp_kernel<1, 2|mlp_rev" -s 3 -c 3 -o /tmp/prof_mlp2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof_mlp2.ncu-rep --page raw --csv > gpurun_out/prof_mlp2_raw.csv 2>/dev/null
ncu -i /tmp/prof_mlp2.ncu-rep --page source --csv 2>/dev/null | python - <<'PY' > gpurun_out/prof_mlp2_stalls.txt
import sys, csv, re, collections
rows = list(csv.reader(sys.stdin))
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}; secs.append(cur)
    elif cur is not None:
        if cur["hdr"] is None: cur["hdr"] = r
        else: cur["rows"].append(r)
seen = set()
for sec in secs:
    if sec["name"] in seen: continue
    seen.add(sec["name"])
    hdr = sec["hdr"]; idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter(); byop = collections.defaultdict(collections.Counter); n = 0
    for r in sec["rows"]:
        if len(r) < len(hdr): continue
        try: ns = int(r[idx["# Samples"]])
        except Exception: continue
        n += ns
        src = r[idx["Source"]].strip()
        op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0].split(".")[0] if src else "?"
        byop[op]["_n"] += ns
        for s_ in stalls:
            v = int(r[idx[s_]] or 0)
            if v: tot[s_] += v; byop[op][s_] += v
    print("=====", sec["name"], "samples", n)
    for k, v in tot.most_common(8): print(f"   {k:26s} {v:9d} {100*v/max(n,1):5.1f}%")
    for op, c in sorted(byop.items(), key=lambda kv: -kv[1]["_n"])[:12]:
        m = c.pop("_n"); print(f"   op {op:10s} {m:9d} {100*m/max(n,1):5.1f}%  {dict(c.most_common(3))}")
PY
ls -la gpurun_out | tail -12; du -sh gpurun_out
