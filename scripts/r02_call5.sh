#!/bin/bash
# Round 2, one GPU: the GPU suite on the trimmed library (new defaults: register-staged C2R, four-CTA 3/2-rule rows,
# y-blocked single-rank layout), then y-blocked vs natural on the single-GPU workloads.
O=gpurun_out/r02_yblock
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
timeout 900 python scripts/ab_single.py --steps 10 --workloads slab1024_f64,slab1024_f64_32,slab512_f64,slab256_f32 > $O/ab_single.jsonl 2> $O/ab_single.txt; cat $O/ab_single.txt
