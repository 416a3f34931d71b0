mkdir -p gpurun_out
timeout -k 10 900 python scripts/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/sweep.log | tail -40
