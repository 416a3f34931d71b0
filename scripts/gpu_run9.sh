mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -k "not full_size" > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout -k 10 300 python bench.py --rows 1250000 --steps 40 --no-cpu-baseline > gpurun_out/bench_1p25m.log 2>&1; echo "1.25m rc=$?"
timeout -k 10 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_10m.log 2>&1; echo "10m rc=$?"
timeout -k 10 600 python scripts/churn_bench.py --rounds 20 > gpurun_out/churn.log 2>&1; echo "churn rc=$?"; tail -3 gpurun_out/churn.log
python - <<'PY'
import json
for f in ("bench_1p25m","bench_10m"):
    try:
        l=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        r=l["roofline"]; print(f, round(l["value"]), round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), l["gpu_launches"], {k:r[k] for k in ("kernel_ms","achieved_tflops","slabs","stages","grid","kprime")})
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.log").read()[-2500:])
PY
