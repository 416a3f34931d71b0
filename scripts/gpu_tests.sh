# full GPU test-suite (what the driver runs at round end)
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/tests_full.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/tests_full.log
