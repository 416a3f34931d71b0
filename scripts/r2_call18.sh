mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_routing.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/r2_tests_4gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_4gpu.log
timeout -k 10 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 40 --warmup 5 > gpurun_out/r2_bench_n4.log 2> gpurun_out/r2_bench_n4.err; echo "bench4 rc=$?"; tail -c 600 gpurun_out/r2_bench_n4.err
tail -1 gpurun_out/r2_bench_n4.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('ms_per_step_one_batch_in_flight'), d['e2e']['value'], d['e2e_one_handle']['value'], d['parity']['ok'], d['roofline']['kernel_ms_per_rank'], d['config']['shard_rows']); print([c['spread'] for c in d['calibration']]); print({n:(round(r['value'],1), r['parity']['ok']) for n,r in d['extra_workloads'].items()}); print(d['parity_failures'])"
