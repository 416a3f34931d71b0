mkdir -p gpurun_out
timeout -k 10 900 python bench.py > gpurun_out/r2_bench_n1.log 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_n1.err
tail -1 gpurun_out/r2_bench_n1.log | cut -c1-3000
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/r2_tests_after_ring.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_after_ring.log
