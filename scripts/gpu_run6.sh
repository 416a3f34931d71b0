mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -k "not full_size" > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/tests.log
timeout -k 10 600 python scripts/churn_bench.py > gpurun_out/churn.log 2>&1; echo "churn rc=$?"; tail -5 gpurun_out/churn.log
timeout -k 10 300 python bench.py --rows 2000000 --k 100 --qbatch 256 --steps 10 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "c4 rc=$?"
timeout -k 10 300 python bench.py --rows 2000000 --k 40 --qbatch 256 --steps 10 --no-cpu-baseline > gpurun_out/bench_k40.log 2>&1; echo "k40 rc=$?"
timeout -k 10 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_10m.log 2>&1; echo "10m rc=$?"
python - <<'PY'
import json
for f in ("bench_c4","bench_k40","bench_10m"):
    try:
        l=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        r=l["roofline"]; print(f, round(l["value"]), round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), {k:r[k] for k in ("kernel_ms","achieved_tflops","achieved_hbm_gbs","slabs","stages","grid","kprime")})
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.log").read()[-1500:])
PY
