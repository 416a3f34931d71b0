mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_routing.py -m gpu -x -q -s > gpurun_out/r2_routing_tests.log 2>&1; echo "routing rc=$?"; tail -5 gpurun_out/r2_routing_tests.log
timeout -k 10 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_routing.py > gpurun_out/r2_tests_full.log 2>&1; echo "full rc=$?"; tail -5 gpurun_out/r2_tests_full.log
timeout -k 10 600 python scripts/anyk_probe.py > gpurun_out/r2_anyk.log 2>&1; echo "anyk rc=$?"; tail -8 gpurun_out/r2_anyk.log
