# round 2 ncu evidence (ONE B200): launch lists + one full capture of the dominant kernel for
#   (a) the headline shape 10 M x 768, Q = 1024   (b) the N = 8 shard size 1.25 M x 768, Q = 1024   (c) Q = 256 at 10 M
#   (d) the select path (k = 500 for 64 queries, 2 M rows)  -- numbers printed under ncu are never bench values
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --extras none"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_10m_q1024.csv $B > gpurun_out/r2_ncu_list_a.log 2>&1; echo "list a rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:fused_knn -s 2 -c 1 -f -o gpurun_out/r2_fused_10m_q1024 $B > gpurun_out/r2_ncu_full_a.log 2>&1; echo "full a rc=$?"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_1p25m_q1024.csv $B --rows 1250000 > gpurun_out/r2_ncu_list_b.log 2>&1; echo "list b rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:fused_knn -s 2 -c 1 -f -o gpurun_out/r2_fused_1p25m_q1024 $B --rows 1250000 > gpurun_out/r2_ncu_full_b.log 2>&1; echo "full b rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:fused_knn -s 2 -c 1 -f -o gpurun_out/r2_fused_10m_q256 $B --qbatch 256 > gpurun_out/r2_ncu_full_c.log 2>&1; echo "full c rc=$?"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_select_2m.csv python scripts/anyk_probe.py --rows 2000000 --reps 1 > gpurun_out/r2_ncu_list_d.log 2>&1; echo "list d rc=$?"
ls -la gpurun_out/*.ncu-rep
