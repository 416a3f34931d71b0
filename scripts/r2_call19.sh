mkdir -p gpurun_out
timeout -k 10 300 python scripts/sweep.py --rows 1250000 --what none --ab certainty=0,1 --q 1024 --rounds 4 --reps 30 > gpurun_out/r2_certainty_ab2.log 2>&1; echo "rc=$?"; grep AB-sample gpurun_out/r2_certainty_ab2.log
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"merge_refine|pack_queries|fused_knn" -c 24 --csv --log-file gpurun_out/r2_launches_1p25m_b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --extras none --rows 1250000 > /dev/null 2>&1; grep merge_refine gpurun_out/r2_launches_1p25m_b.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final2.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_final2.log
