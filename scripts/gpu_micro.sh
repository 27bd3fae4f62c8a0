#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_passes.py -m gpu -x -q 2>&1 | tail -3
for v in 3 5 6 7 0; do
  B200FFT_VARIANT=$v timeout 300 python scripts/microbench_strided.py 1024 d 2>&1 | tee -a gpurun_out/micro_strided.log
done
B200FFT_VARIANT=3 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_1024_v3.json 2> gpurun_out/bench_1024_v3.err
python scripts/show_passes.py gpurun_out/bench_1024_v3.json
B200FFT_VARIANT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --workload slab1024_f64_32 > gpurun_out/bench_1536_v0.json 2> gpurun_out/bench_1536_v0.err
python scripts/show_passes.py gpurun_out/bench_1536_v0.json
