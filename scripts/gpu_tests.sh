# full GPU test-suite + smoke (what the driver runs at round end)
mkdir -p gpurun_out
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/tests_full.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_full.log
