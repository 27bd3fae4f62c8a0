#!/bin/bash
# Round 2, call 1 (one GPU): the whole GPU suite with its log, then the A/B of every opt-in kernel / schedule
# in one process, then the far-stride microbenchmark and the ncu launch list of the default bench.
O=gpurun_out/r02_single
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -30 $O/pytest_gpu.log
timeout 1500 python scripts/ab_single.py --steps 5 > $O/ab_single.jsonl 2> $O/ab_single.txt; tail -120 $O/ab_single.txt
for v in 0 20 21 23; do
  B200FFT_VARIANT=$v timeout 200 python scripts/microbench_strided.py 1024 d 2>&1 | tee -a $O/micro_strided.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft_ -c 60 --csv --log-file $O/launches_1024.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launch.log 2>&1
tail -3 $O/ncu_launch.log
ls -la $O
