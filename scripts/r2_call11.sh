mkdir -p gpurun_out
timeout -k 10 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2_bench_n2.log 2> gpurun_out/r2_bench_n2.err; echo "bench2 rc=$?"; tail -c 800 gpurun_out/r2_bench_n2.err
tail -1 gpurun_out/r2_bench_n2.log | cut -c1-400
