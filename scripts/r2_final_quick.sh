mkdir -p gpurun_out
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests_final.log
timeout -k 10 300 python bench.py --extras c5 --no-cpu-baseline --steps 10 > gpurun_out/r2_bench_quick.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_quick.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['parity']['ok'], d['extra_workloads']['c5']['host_input_gbs'], d['extra_workloads']['c5']['parity']['ok'])"
