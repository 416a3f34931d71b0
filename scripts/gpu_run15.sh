mkdir -p gpurun_out
for n in 8 4; do
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 40 --warmup 5 > gpurun_out/bench_g$n.log 2>&1; echo "bench g$n rc=$?"
done
python - <<'PY'
import json
for f in ("bench_g8","bench_g4"):
    try:
        l=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        r=l["roofline"]; print(f, round(l["value"]), round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), l["gpu_launches"], {k:r[k] for k in ("kernel_ms","achieved_tflops","slabs","stages","grid","kprime")}, l["clocks"])
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.log").read()[-3000:])
PY
