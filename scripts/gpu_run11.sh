mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cta_pair" > gpurun_out/tests_pair.log 2>&1; echo "pair tests rc=$?"; tail -15 gpurun_out/tests_pair.log
timeout -k 10 600 python -m pytest tests -m gpu -q -k "not full_size and not cta_pair" > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
for cp in 1 2; do PWKNN_CTA_PAIR=$cp timeout -k 10 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_10m_cp$cp.log 2>&1; echo "bench cp$cp rc=$?"; done
python - <<'PY'
import json
for f in ("bench_10m_cp1","bench_10m_cp2"):
    try:
        l=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        r=l["roofline"]; print(f, round(l["value"]), round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), {k:r[k] for k in ("kernel_ms","achieved_tflops","slabs","stages","grid","kprime")}, l["clocks"])
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.log").read()[-2500:])
PY
