mkdir -p gpurun_out
timeout -k 10 300 python scripts/sweep.py --rows 1250000 --what none --ab certainty=0,1 --q 1024 --rounds 4 --reps 30 > gpurun_out/r2_certainty_ab.log 2>&1; echo "rc=$?"; grep AB gpurun_out/r2_certainty_ab.log
