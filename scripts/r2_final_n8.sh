mkdir -p gpurun_out
timeout -k 10 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r2_bench_final_n8.log 2> gpurun_out/r2_bench_final_n8.err; echo "bench8 rc=$?"; tail -c 500 gpurun_out/r2_bench_final_n8.err
tail -1 gpurun_out/r2_bench_final_n8.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_one_handle']['value'], d['parity']['ok'], d['roofline']['kernel_ms_per_rank'], d['roofline']['frac']); print([c['spread'] for c in d['calibration']], d['config']['shard_rows']); print({n:(round(r['value'],1), r['parity']['ok']) for n,r in d['extra_workloads'].items()}); print(d['parity_failures'])"
