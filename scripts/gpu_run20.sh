mkdir -p gpurun_out
timeout -k 10 600 python scripts/sweep.py --ab lead_tiles=0,1,2,4,8 --reps 10 --rounds 8 > gpurun_out/ab_lead.log 2>&1; echo "ab rc=$?"; grep AB gpurun_out/ab_lead.log
for lt in 1 2 4; do
PWKNN_LEAD=$lt timeout -k 10 600 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fused_knn -s 2 -c 1 --csv --log-file gpurun_out/ncu_lead$lt.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_lead$lt.log 2>&1
echo -n "lead $lt: "; grep -E "fused_knn" gpurun_out/ncu_lead$lt.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr '\n' ' '; echo
done
