# Regenerates the evidence under profiles/ (run through gpurun on ONE B200):
#   launch list of a step, full ncu capture of the dominant kernel, plus the complete GPU test-suite.
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q > gpurun_out/tests_full.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_full.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_knn|merge_refine|pack_queries" -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "nculist rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:fused_knn -s 2 -c 1 -f -o gpurun_out/prof_fused_final python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncufull rc=$?"
timeout -k 10 600 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final.log | cut -c1-1800
