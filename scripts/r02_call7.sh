#!/bin/bash
# Round 2, one GPU: GPU suite (new: NS kernels, eager + CUDA graph), the default bench line, ncu launch list and
# full-set captures of the six passes of the headline transform and of the 3/2-rule one.
O=gpurun_out/r02_final1
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_slab1024_f64.json 2> $O/bench_slab1024_f64.err; python scripts/show_passes.py $O/bench_slab1024_f64.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --workload slab1024_f64_32 > $O/bench_slab1024_f64_32.json 2> $O/bench_32.err; python scripts/show_passes.py $O/bench_slab1024_f64_32.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft_ -c 80 --csv --log-file $O/launches_1024.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launch.log 2>&1
for w in slab1024_f64 slab1024_f64_32; do
  # skip the parity transform (3 kernels) and three warm-up round trips (18), capture one round trip (6)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_ -s 21 -c 6 -o $O/prof_$w -f \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload $w > $O/ncu_full_$w.log 2>&1
  ncu -i $O/prof_$w.ncu-rep --page raw --csv > $O/prof_$w.raw.csv 2>/dev/null
done
ls -la $O
