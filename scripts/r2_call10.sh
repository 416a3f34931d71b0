mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout -k 10 900 python -m pytest tests/test_sharded_nccl.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/r2_tests_2gpu.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests_2gpu.log
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2_bench_n2.log 2> gpurun_out/r2_bench_n2.err; echo "bench2 rc=$?"; tail -c 1200 gpurun_out/r2_bench_n2.err
tail -1 gpurun_out/r2_bench_n2.log | cut -c1-600
