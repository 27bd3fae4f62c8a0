#!/bin/bash
# Round-2 opener, ONE GPU (gpurun --timeout 1500 -- 'bash scripts/gpu_round2_single.sh'):
# device runs of the opt-in kernels written blind at the end of round 1, then A/B timings.
# Everything goes to gpurun_out/r02_single/.
O=gpurun_out/r02_single
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
# 1. parity first: the default path, then the experimental variants (subprocess, xfail until verified)
timeout 1200 python -m pytest tests -m gpu -q -rxX > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
# 1b. compute-sanitizer over small launches of every kernel family and variant (shared-memory races, barrier
#     misuse, out-of-bounds): the checks the CPU emulator cannot make
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tests/gpu_sanitize_worker.py > $O/sanitize_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_WORKER_OK|variant .* ok|Error" $O/sanitize_$tool.log | tail -12
done
# 2a. everything below in one process first (one import, one set of arrays): the quick overall picture
timeout 1500 python scripts/ab_single.py --steps 5 > $O/ab_single.jsonl 2> $O/ab_single.txt; tail -80 $O/ab_single.txt
# 2a'. the tuner's own view (same candidates through mpifft4py_b200.tune.autotune)
for w in slab1024_f64 slab1024_f64_32; do
  timeout 900 python scripts/tune_single.py $w patient > $O/tune_$w.json 2> $O/tune_$w.err; tail -3 $O/tune_$w.err; grep chosen $O/tune_$w.json
done
# 2. cluster strided pass (variant 20: far launches only), per-row barriers in the row kernels (30),
#    register-staged C2R (31)
#    against the default, plain and 3/2-rule
for v in 0 20 30 31 32 33 34 35; do
  for w in slab1024_f64 slab1024_f64_32; do
    B200FFT_VARIANT=$v timeout 400 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --workload $w \
        > $O/bench_${w}_v$v.json 2> $O/bench_${w}_v$v.err
    echo "== $w variant $v"; python scripts/show_passes.py $O/bench_${w}_v$v.json; tail -2 $O/bench_${w}_v$v.err
  done
done
# 2b. L2 blocking of the z / y passes: groups of x planes (0 = off)
for g in 2 4 6 8 12 16 -2 -3 -4 -6 f2 f3 f4 f6 f8 f12; do   # -g: two streams; fg: one fused persistent kernel
  for w in slab1024_f64 slab1024_f64_32; do
    case $g in -*) mode=2;; f*) mode=3;; *) mode=1;; esac
    B200FFT_L2_MODE=$mode B200FFT_L2_PLANES=${g#[-f]} timeout 400 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --workload $w \
        > $O/bench_${w}_l2_$g.json 2> $O/bench_${w}_l2_$g.err
    echo "== $w l2_planes $g"; python scripts/show_passes.py $O/bench_${w}_l2_$g.json; tail -2 $O/bench_${w}_l2_$g.err
  done
done
# 2c. which side of a far-strided pass pays: loads or stores (default kernels, cluster kernels 128 / 64 B)
for v in 0 20 21 23; do
  B200FFT_VARIANT=$v timeout 300 python scripts/microbench_strided.py 1024 d 2>&1 | tee -a $O/micro_strided.log
done
# 3. ncu: launch list of the default bench, full capture of the x pass in both variants
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft_ -c 60 --csv --log-file $O/launches_1024.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launch.log 2>&1
for v in 0 20; do
  B200FFT_VARIANT=$v timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_ -s 18 -c 6 -o $O/prof_1024_v$v -f \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_v$v.log 2>&1
done
ls -la $O
