mkdir -p gpurun_out
timeout -k 10 600 python bench.py --workload c5 > gpurun_out/r2_c5.log 2> gpurun_out/r2_c5.err; echo "c5 rc=$?"; tail -c 600 gpurun_out/r2_c5.err; tail -1 gpurun_out/r2_c5.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())['detail']; print(json.dumps({k:d[k] for k in ('value','host_input_gbs','ms_per_round','qps_static')}))"
timeout -k 10 500 python scripts/sweep.py --what none --ab margin=6,22 --q 1024 --rounds 4 --reps 8 > gpurun_out/r2_margin_ab2.log 2>&1; echo "margin rc=$?"; tail -3 gpurun_out/r2_margin_ab2.log
