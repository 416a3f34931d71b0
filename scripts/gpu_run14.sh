mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x -k "not full_size" > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/tests.log
timeout -k 10 300 python bench.py --rows 2000000 --k 100 --qbatch 256 --steps 10 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "c4 rc=$?"
timeout -k 10 300 python bench.py --rows 10000000 --k 100 --qbatch 1024 --steps 10 --no-cpu-baseline > gpurun_out/bench_k100.log 2>&1; echo "k100 rc=$?"
timeout -k 10 300 python scripts/sweep.py --rows 60000 --what q --reps 10 > gpurun_out/sweep_l2.log 2>&1; echo "sweep rc=$?"; grep -E '"Q": (1024|4096)' gpurun_out/sweep_l2.log | cut -c1-200
python - <<'PY'
import json
for f in ("bench_c4","bench_k100"):
    try:
        l=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        r=l["roofline"]; print(f, round(l["value"]), round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), l["gpu_launches"], {k:r[k] for k in ("kernel_ms","achieved_tflops","slabs","stages","grid","kprime")})
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.log").read()[-2500:])
PY
