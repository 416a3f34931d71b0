mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -k "not full_size" > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/tests.log
timeout -k 10 300 python bench.py --rows 1250000 --steps 40 --no-cpu-baseline > gpurun_out/bench_1p25m.log 2>&1; echo "1.25m rc=$?"
timeout -k 10 300 python bench.py --rows 2000000 --k 100 --qbatch 256 --steps 10 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "c4 rc=$?"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_knn|merge_refine|pack_queries" -c 30 --csv --log-file gpurun_out/launches_1p25m.csv python bench.py --rows 1250000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "nculist rc=$?"
python - <<'PY'
import json,csv
for f in ("bench_1p25m","bench_c4"):
    try:
        l=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        r=l["roofline"]; print(f, round(l["value"]), round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), l["gpu_launches"], {k:r[k] for k in ("kernel_ms","achieved_tflops","slabs","stages","grid","kprime")})
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.log").read()[-2500:])
rows=[r for r in csv.reader(open('gpurun_out/launches_1p25m.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
from collections import defaultdict
agg=defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki].split('(')[0][-40:]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in agg.items(): print(k, len(v), sum(v)/len(v))
PY
