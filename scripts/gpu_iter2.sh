#!/bin/bash
# tuning iteration: pass tests, bench both headline workloads, ncu full capture of both (CSV exports only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_passes.py tests/test_gpu_transforms.py -m gpu -x -q 2>&1 | tail -4
for w in slab1024_f64 slab1024_f64_32 slab256_f32 slab512_f64; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python scripts/show_passes.py gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err
done
for w in slab1024_f64 slab1024_f64_32; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 18 -c 6 -o /tmp/prof_$w -f \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload $w > gpurun_out/ncu_$w.log 2>&1
  ncu -i /tmp/prof_$w.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$w.csv 2>/dev/null
  ncu -i /tmp/prof_$w.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_src_$w.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
