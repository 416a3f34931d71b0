mkdir -p gpurun_out
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_final.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_knn|merge_refine|pack_queries" -c 40 --csv --log-file gpurun_out/r2_launches_10m_q1024.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/r2_ncu_list_a.log 2>&1; echo "list a rc=$?"
timeout -k 10 900 python bench.py > gpurun_out/r2_bench_final_n1.log 2> gpurun_out/r2_bench_final_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_final_n1.err; tail -1 gpurun_out/r2_bench_final_n1.log | cut -c1-400
