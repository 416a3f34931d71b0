for t in 1 4; do PWKNN_HOST_THREADS=$t timeout -k 10 200 python scripts/host_stage_probe.py 2>&1 | tail -2; done
