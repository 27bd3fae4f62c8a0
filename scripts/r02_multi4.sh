#!/bin/bash
# Round 2, 4 GPUs: parity worker (every class / alignment / communication / dealias mode / golden / known answers,
# default copy-engine transport + NCCL for one object per class), the worker's transport modes, then the default
# configuration of each BASELINE workload that fits 4 ranks.
N=4
O=gpurun_out/r02_multi_4
mkdir -p $O
port() { echo $((29500 + RANDOM % 2000)); }
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) "${@:2}"; }
tr 200 tests/gpu_dist_worker.py > $O/parity_worker.log 2>&1
echo "parity worker rc=$? ($(grep -c GPU_WORKER_OK $O/parity_worker.log) of $N ranks ok)" | tee -a $O/summary.txt
grep "^\[" $O/parity_worker.log | tail -4
for mode in "store x" "p2p kz" "nccl pencil-chunks" "p2p pencil-chunks"; do
  tag=$(echo $mode | tr ' ' '_')
  tr 120 tests/gpu_dist_worker.py --transport $mode > $O/parity_$tag.log 2>&1
  echo "worker --transport $mode rc=$? ($(grep -c GPU_WORKER_OK $O/parity_$tag.log) of $N ranks ok)" | tee -a $O/summary.txt
done
tr 150 scripts/ab_multi.py --steps 10 --workloads slab1024_f64,slab1024_f64_32,pencilX1024_f64,pencilY2048_f32,line16384_f32 --configs default,nccl_c1 \
    > $O/ab_defaults.jsonl 2> $O/ab_defaults.txt
echo "ab defaults rc=$?" | tee -a $O/summary.txt; grep "^slab\|^pencil\|^line" $O/ab_defaults.txt
