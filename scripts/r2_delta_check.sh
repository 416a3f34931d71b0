mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_routing.py tests/test_gpu_parity.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/r2_tests_delta4.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_delta4.log
for i in 1 2; do timeout -k 10 600 python bench.py --workload c5 > gpurun_out/r2_c5.log 2> gpurun_out/r2_c5.err; echo "c5 rc=$?"; tail -c 300 gpurun_out/r2_c5.err; tail -1 gpurun_out/r2_c5.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())['detail']; print(json.dumps({k:d[k] for k in ('value','host_input_gbs','ms_per_round','qps_static')}), d['parity']['ok'])"; done
