# A/B of the sibling lead limit at the 8-GPU shard size (one GPU) + a regression subset: scripts/sweep.py --ab
mkdir -p gpurun_out
timeout -k 10 200 python scripts/sweep.py --rows 1250000 --what none --ab lead_tiles=4,1 --q 1024 --rounds 6 --reps 30 > gpurun_out/r2_lead_small_after.log 2>&1; grep "^AB " gpurun_out/r2_lead_small_after.log | cut -c1-110
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab or cta_pair or properties or kat or lead" > gpurun_out/r2_tests_lead.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests_lead.log
