mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_routing.py tests/test_gpu_parity.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/r2_tests_delta2.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_delta2.log
timeout -k 10 600 python bench.py --workload c5 > gpurun_out/r2_c5.log 2> gpurun_out/r2_c5.err; echo "c5 rc=$?"; tail -c 600 gpurun_out/r2_c5.err; tail -1 gpurun_out/r2_c5.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())['detail']; print(json.dumps({k:d[k] for k in ('value','host_input_gbs','ms_per_round','qps_static','parity')}))"
timeout -k 10 600 python bench.py --extras c2_raw_f32 --no-cpu-baseline --steps 10 > gpurun_out/r2_rawf32.log 2> gpurun_out/r2_rawf32.err; echo "raw rc=$?"; tail -1 gpurun_out/r2_rawf32.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())['extra_workloads']['c2_raw_f32']; print(json.dumps({k:d[k] for k in ('value','ms_per_step','kernel_ms','roofline_frac','kprime')})); print(d['e2e'])"
