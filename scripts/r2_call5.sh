mkdir -p gpurun_out
timeout -k 10 900 python bench.py > gpurun_out/r2_bench_n1.log 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_n1.err
tail -1 gpurun_out/r2_bench_n1.log | cut -c1-1200
