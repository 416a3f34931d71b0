mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/sharded_nccl_check.py > gpurun_out/sharded_nccl.log 2>&1; echo "sharded rc=$?"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_g2.log 2>&1; echo "bench g2 rc=$?"
timeout -k 10 300 python bench.py --impl reference --gpus 1 --steps 6 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
for f in tests sharded_nccl bench_g2 bench_ref; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-3000; done
