mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x -k "not full_size" > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/tests.log
timeout -k 10 600 python bench.py --no-cpu-baseline > gpurun_out/bench_10m.log 2>&1; echo "bench rc=$?"
timeout -k 10 600 python scripts/sweep.py --what q --reps 20 > gpurun_out/sweep_q.log 2>&1; echo "sweep rc=$?"; grep -E '"Q": (128|256|512|1024|4096),' gpurun_out/sweep_q.log | cut -c1-175
python - <<'PY'
import json
for f in ("bench_10m",):
    l=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
    r=l["roofline"]; print(f, round(l["value"]), round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), l["gpu_launches"], {k:r[k] for k in ("kernel_ms","achieved_tflops","frac","slabs","stages","grid","kprime")}, l["clocks"])
PY
