mkdir -p gpurun_out
for cp in 2 1; do
PWKNN_CTA_PAIR=$cp timeout -k 10 600 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,smsp__inst_executed.sum,dram__bytes_read.sum --clock-control none -k regex:fused_knn -s 2 -c 2 --csv --log-file gpurun_out/ncu_quick_cp$cp.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_quick_cp$cp.log 2>&1; echo "ncu cp$cp rc=$?"
grep -E "fused_knn" gpurun_out/ncu_quick_cp$cp.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-200 | tail -6
done
