#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_passes.py tests/test_gpu_transforms.py -m gpu -x -q 2>&1 | tail -3
for v in 0 3; do
  echo "=== variant $v"
  B200FFT_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --workload slab1024_f64 > gpurun_out/bench_1024_v$v.json 2> gpurun_out/bench_1024_v$v.err
  python scripts/show_passes.py gpurun_out/bench_1024_v$v.json; tail -2 gpurun_out/bench_1024_v$v.err
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --workload slab1024_f64_32 > gpurun_out/bench_1536.json 2> gpurun_out/bench_1536.err
python scripts/show_passes.py gpurun_out/bench_1536.json; tail -2 gpurun_out/bench_1536.err
