# round 2, call 1: noise model inputs, sanitizer logs, k'=32 cost, f32-store headline baseline
mkdir -p gpurun_out
timeout -k 10 300 python scripts/noise_probe.py > gpurun_out/r2_noise.log 2>&1; echo "noise rc=$?"
timeout -k 10 420 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta_pair_variant and 5000 or reference_kat_gpu" > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?"
timeout -k 10 300 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta_pair_variant and 5000 or reference_kat_gpu" > gpurun_out/r2_synccheck.log 2>&1; echo "synccheck rc=$?"
timeout -k 10 400 python scripts/sweep.py --what none --ab margin=6,22 --q 1024 --rounds 4 --reps 8 > gpurun_out/r2_margin_ab.log 2>&1; echo "margin rc=$?"
timeout -k 10 500 python bench.py --exact-rerank f32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_f32store.log 2>&1; echo "f32 rc=$?"
tail -3 gpurun_out/r2_noise.log; tail -3 gpurun_out/r2_racecheck.log; tail -3 gpurun_out/r2_synccheck.log; tail -3 gpurun_out/r2_margin_ab.log; tail -1 gpurun_out/r2_bench_f32store.log | cut -c1-600
