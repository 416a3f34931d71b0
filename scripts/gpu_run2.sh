mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
timeout -k 10 300 python bench.py --rows 2000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2m.log 2>&1; echo "bench2m rc=$?"
timeout -k 10 600 python bench.py > gpurun_out/bench_10m.log 2>&1; echo "bench10m rc=$?"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_knn|merge_refine|pack_queries" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "nculist rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:fused_knn -s 2 -c 1 -f -o gpurun_out/prof_fused python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncufull rc=$?"
for f in tests bench_2m bench_10m; do echo "=== $f"; tail -n 8 gpurun_out/$f.log; done
tail -n 5 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
