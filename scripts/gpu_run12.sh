mkdir -p gpurun_out
timeout -k 10 600 python scripts/sweep.py --what q > gpurun_out/sweep_pair.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/sweep_pair.log | cut -c1-230
timeout -k 10 600 python bench.py > gpurun_out/bench_10m.log 2>&1; echo "bench rc=$?"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_knn|merge_refine|pack_queries" -c 60 --csv --log-file gpurun_out/launches_pair.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "nculist rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:fused_knn -s 2 -c 1 -f -o gpurun_out/prof_fused_pair python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncufull rc=$?"
tail -1 gpurun_out/bench_10m.log | cut -c1-2500
