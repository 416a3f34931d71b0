mkdir -p gpurun_out
# L2-resident corpus: is the remaining ~13% tensor idle time HBM first-touch latency?
timeout -k 10 300 python scripts/sweep.py --rows 60000 --what q --reps 10 > gpurun_out/sweep_l2.log 2>&1; echo "sweep rc=$?"; grep -E '"Q": (1024|2048|4096)' gpurun_out/sweep_l2.log | cut -c1-200
timeout -k 10 300 python scripts/sweep.py --rows 2500000 --what q --reps 10 > gpurun_out/sweep_2p5m.log 2>&1; echo "sweep rc=$?"; grep -E '"Q": (1024|4096)' gpurun_out/sweep_2p5m.log | cut -c1-200
