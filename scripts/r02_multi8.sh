#!/bin/bash
# Round 2, 8 GPUs (gpurun --gpus 8 -- 'bash scripts/r02_multi8.sh'): parity of every class / alignment / communication /
# dealias mode / golden at 8 ranks, then the transport x pipeline A/B (forward parity inside every case) on the
# BASELINE workloads.  Tight timeouts: 8-GPU minutes cost eight.
N=${1:-8}
O=gpurun_out/r02_multi_$N
mkdir -p $O
port() { echo $((29500 + RANDOM % 2000)); }
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) "${@:2}"; }
tr 240 tests/gpu_dist_worker.py > $O/parity_worker.log 2>&1
echo "parity worker rc=$? ($(grep -c GPU_WORKER_OK $O/parity_worker.log) of $N ranks ok)" | tee -a $O/summary.txt
ab() {  # tag timeout workloads configs
  tr $2 scripts/ab_multi.py --steps 10 --workloads $3 --configs $4 > $O/ab_$1.jsonl 2> $O/ab_$1.txt
  echo "ab $1 rc=$?" | tee -a $O/summary.txt; grep "^slab\|^pencil\|^line" $O/ab_$1.txt | tail -40
}
ab slab 150 slab1024_f64,slab1024_f64_32 default,p2p_c1,p2p_c2,p2p_c4,p2p_c8,p2p_c2_cs,p2p_c4_cs,p2p_c8_cs,nccl_c1,nccl_c2
ab slab_kz 120 slab1024_f64,slab1024_f64_32 p2p_kz2,p2p_kz4,p2p_kz4_cs,p2p_kz8_cs,nccl_kz2
ab other_nccl 150 pencilX1024_f64,pencilX512_f64,pencilY2048_f32,line16384_f32,slab256_f32 default,nccl_c2,nccl_c4
ab other_p2p 120 pencilX1024_f64,pencilX512_f64,pencilY2048_f32,line16384_f32,slab256_f32 p2p_c1,p2p_c2,p2p_c4
ab store 100 slab1024_f64,pencilX1024_f64 store_c1,store_kz4
ls -la $O
