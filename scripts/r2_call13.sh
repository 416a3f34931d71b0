mkdir -p gpurun_out
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 40 --warmup 5 --extras none > gpurun_out/r2_bench_n2b.log 2> gpurun_out/r2_bench_n2b.err; echo "bench2 rc=$?"; tail -c 600 gpurun_out/r2_bench_n2b.err
tail -1 gpurun_out/r2_bench_n2b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('ms_per_step_one_batch_in_flight'), d['e2e']['value'], d['parity']['ok'], d['roofline']['kernel_ms_per_rank'], d['config']['shard_rows']); print(d['calibration'])"
