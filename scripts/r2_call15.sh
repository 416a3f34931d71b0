mkdir -p gpurun_out
timeout -k 10 600 python scripts/cta_times_probe.py > gpurun_out/r2_cta_times.log 2>&1; echo "cta rc=$?"; cat gpurun_out/r2_cta_times.log | cut -c1-700
