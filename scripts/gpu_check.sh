#!/bin/bash
# One gpurun call: smoke, GPU tests, bench, ncu launch list + one full capture.  Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_1024.json; tail -5 gpurun_out/bench_1024.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload slab256_f32 --no-cpu-baseline > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload slab1024_f64_32 --no-cpu-baseline --no-e2e > gpurun_out/bench_1024_32.json 2> gpurun_out/bench_1024_32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft_kernel -c 60 --csv --log-file gpurun_out/launches_1024.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 18 -c 6 -o gpurun_out/prof_1024 -f \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
