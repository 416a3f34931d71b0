mkdir -p gpurun_out
timeout -k 10 400 python bench.py --impl reference --steps 3 > gpurun_out/r2_bench_ref2.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/r2_bench_ref2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['cpu_baseline']['linearity'], d['cpu_baseline']['all_cores'])"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
