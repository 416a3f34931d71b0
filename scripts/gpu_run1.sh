mkdir -p gpurun_out; nvidia-smi > gpurun_out/nvsmi.txt 2>&1
timeout -k 10 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "create_add or exact_path" > gpurun_out/t1.log 2>&1; echo "t1 rc=$?"
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tcgen05_scores" > gpurun_out/t2.log 2>&1; echo "t2 rc=$?"
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not tcgen05_scores and not create_add and not exact_path" > gpurun_out/t3.log 2>&1; echo "t3 rc=$?"
timeout -k 10 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout -k 10 500 python bench.py --rows 2000000 --steps 10 --warmup 3 > gpurun_out/bench_2m.log 2>&1; echo "bench rc=$?"
for f in t1 t2 t3 smoke bench_2m; do echo "=== $f"; tail -n 25 gpurun_out/$f.log; done
