#!/bin/bash
# N GPUs: parity worker with the flag kernel, then slab 1024^3 with the flag kernel (default) against the per-peer DMA flags
N=${1:-2}
O=gpurun_out/r02_flags_$N
mkdir -p $O
port() { echo $((29500 + RANDOM % 2000)); }
trun() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) "${@:2}"; }
trun 150 tests/gpu_dist_worker.py > $O/parity_worker.log 2>&1
echo "parity worker rc=$? ($(grep -c GPU_WORKER_OK $O/parity_worker.log) of $N ranks ok)" | tee -a $O/summary.txt
trun 120 scripts/ab_multi.py --steps 20 --workloads slab1024_f64 --configs default,p2p_c4,p2p_kz4 > $O/ab_kernel.jsonl 2> $O/ab_kernel.txt
grep "^slab" $O/ab_kernel.txt | sed 's/^/flag kernel: /'
B200FFT_FLAG_DMA=1 trun 120 scripts/ab_multi.py --steps 20 --workloads slab1024_f64 --configs default,p2p_c4,p2p_kz4 > $O/ab_dma.jsonl 2> $O/ab_dma.txt
grep "^slab" $O/ab_dma.txt | sed 's/^/flag DMA:    /'
