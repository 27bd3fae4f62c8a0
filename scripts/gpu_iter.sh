#!/bin/bash
# One 1-GPU gpurun call per tuning iteration: GPU tests, bench per plan variant, ncu launch list + full capture.
# usage: gpu_iter.sh "<variants for 1024>" "<variants for 1536>" [skiptests]
V1=${1:-0}; V2=${2:-0}; SKIP=${3:-}
mkdir -p gpurun_out
if [ -z "$SKIP" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
  tail -8 gpurun_out/pytest_gpu.log
fi
for v in $V1; do
  B200FFT_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_1024_v$v.json 2> gpurun_out/bench_1024_v$v.err
  echo "== 1024 variant $v rc=$?"; python scripts/show_passes.py gpurun_out/bench_1024_v$v.json; tail -2 gpurun_out/bench_1024_v$v.err
done
for v in $V2; do
  B200FFT_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --workload slab1024_f64_32 > gpurun_out/bench_1536_v$v.json 2> gpurun_out/bench_1536_v$v.err
  echo "== 1536 variant $v rc=$?"; python scripts/show_passes.py gpurun_out/bench_1536_v$v.json; tail -2 gpurun_out/bench_1536_v$v.err
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --workload slab256_f32 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err
echo "== slab256_f32"; python scripts/show_passes.py gpurun_out/bench_256.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft_kernel -c 60 --csv --log-file gpurun_out/launches_1024.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 18 -c 6 -o gpurun_out/prof_1024 -f \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20
