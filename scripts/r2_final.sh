# final single-GPU regression: smoke, the complete -m gpu suite, the default bench (what the driver runs)
mkdir -p gpurun_out
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests_final.log
timeout -k 10 900 python bench.py > gpurun_out/r2_bench_final_n1.log 2> gpurun_out/r2_bench_final_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_final_n1.err; tail -1 gpurun_out/r2_bench_final_n1.log | cut -c1-300
