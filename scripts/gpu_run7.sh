mkdir -p gpurun_out
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/sharded_nccl_check.py > gpurun_out/sharded_nccl.log 2>&1; echo "sharded rc=$?"; tail -4 gpurun_out/sharded_nccl.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_g2.log 2>&1; echo "bench g2 rc=$?"
python - <<'PY'
import json
for f in ("bench_g2",):
    try:
        l=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        r=l["roofline"]; print(f, round(l["value"]), round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), l["gpu_launches"], {k:r[k] for k in ("kernel_ms","achieved_tflops","slabs","stages","grid","kprime")}, l["clocks"])
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.log").read()[-2500:])
PY
