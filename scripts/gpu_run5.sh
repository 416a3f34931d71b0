mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -k "not full_size" > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/tests.log
timeout -k 10 600 python scripts/churn_bench.py > gpurun_out/churn.log 2>&1; echo "churn rc=$?"; tail -5 gpurun_out/churn.log
timeout -k 10 300 python bench.py --rows 2000000 --k 100 --qbatch 256 --steps 10 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "c4 rc=$?"; tail -1 gpurun_out/bench_c4.log | cut -c1-600
