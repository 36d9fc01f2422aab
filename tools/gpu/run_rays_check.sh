mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rays.py tests/test_gpu_render.py tests/test_gpu_extra.py tests/test_gpu_extract.py -q -m gpu 2>&1 | grep -E "^E   |passed|failed|FAILED" | cut -c1-250 | head
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > gpurun_out/ncu_launches.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_train.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
data=[(r[kn],float(r[mv].replace(',',''))) for r in rows[hi+1:] if len(r)>mv and r[mv]]
idx=[i for i,d in enumerate(data) if 'coarse_z' in d[0]]
step=data[idx[0]:idx[1]]
print([ (n.split('(')[0][:28], round(v/1e3,1)) for n,v in step if 'upsample' in n or 'render_core' in n or 'render_prep' in n or 'reduce' in n])
PY
