mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests -m gpu -q -x -k "cta_pair or random_parity or slab_count or mid_k" > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout -k 10 600 python scripts/sweep.py --what lead > gpurun_out/sweep_lead.log 2>&1; echo "sweep rc=$?"; cut -c1-150 gpurun_out/sweep_lead.log
for lt in 0 8; do
PWKNN_LEAD=$lt timeout -k 10 600 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg,gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fused_knn -s 2 -c 1 --csv --log-file gpurun_out/ncu_lead$lt.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_lead$lt.log 2>&1; echo "ncu lead$lt rc=$?"
grep -E "fused_knn" gpurun_out/ncu_lead$lt.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr '\n' ' '; echo
done
