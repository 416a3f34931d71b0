mkdir -p gpurun_out
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 40 --warmup 5 --extras none > gpurun_out/r2_bench_n2c.log 2> gpurun_out/r2_bench_n2c.err; echo "bench2 rc=$?"; tail -c 500 gpurun_out/r2_bench_n2c.err
tail -1 gpurun_out/r2_bench_n2c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_one_handle']['value'], d['parity']['ok'], d['roofline']['kernel_ms_per_rank'], d['config']['shard_rows']); print(d['calibration'])"
