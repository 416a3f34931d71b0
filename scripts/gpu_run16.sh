mkdir -p gpurun_out
timeout -k 10 600 python scripts/sweep.py --what diag > gpurun_out/sweep_diag.log 2>&1; echo "sweep rc=$?"; cut -c1-160 gpurun_out/sweep_diag.log
