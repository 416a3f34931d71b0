mkdir -p gpurun_out
timeout -k 10 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r2_bench_n8.log 2> gpurun_out/r2_bench_n8.err; echo "bench8 rc=$?"; tail -c 800 gpurun_out/r2_bench_n8.err
tail -1 gpurun_out/r2_bench_n8.log | cut -c1-300
timeout -k 10 300 python bench.py --impl reference --steps 3 > gpurun_out/r2_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/r2_bench_ref.log | cut -c1-300
