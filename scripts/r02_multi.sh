#!/bin/bash
# Round 2, N GPUs (gpurun --gpus N -- 'bash scripts/r02_multi.sh N'): the parity worker (every class / alignment /
# communication / dealias mode / golden at this rank count), then the A/B of transports x pipelines in a few
# process groups (a protocol mistake in an opt-in mode hangs only its own group until the timeout).
N=${1:-2}
O=gpurun_out/r02_multi_$N
mkdir -p $O
port() { echo $((29500 + RANDOM % 2000)); }
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) "${@:2}"; }
tr 600 tests/gpu_dist_worker.py > $O/parity_worker.log 2>&1
echo "parity worker rc=$? ($(grep -c GPU_WORKER_OK $O/parity_worker.log) of $N ranks ok)" | tee -a $O/summary.txt
tail -5 $O/parity_worker.log
ab() {  # tag timeout workloads configs
  tr $2 scripts/ab_multi.py --steps 10 --workloads $3 --configs $4 > $O/ab_$1.jsonl 2> $O/ab_$1.txt
  echo "ab $1 rc=$?" | tee -a $O/summary.txt; grep -v "^W\|^\[" $O/ab_$1.txt | tail -40
}
ab slab_safe 240 slab1024_f64,slab1024_f64_32 default,nccl_c1,nccl_c2,p2p_c1,p2p_c2,p2p_c4,p2p_c8
ab slab_cs 150 slab1024_f64 p2p_c4_cs,p2p_c8_cs
ab slab_store 150 slab1024_f64,slab1024_f64_32 store_c1,store_c2,store_c4
ab slab_kz 200 slab1024_f64,slab1024_f64_32 nccl_kz2,p2p_kz2,p2p_kz4,p2p_kz4_cs,p2p_kz8_cs
ab slab_store_kz 150 slab1024_f64,slab1024_f64_32 store_kz2,store_kz4,store_kz8
ab line 200 line16384_f32 default,nccl_c2,nccl_c4,p2p_c1,p2p_c2,p2p_c4,store_c1,store_c2
if [ "$N" -ge 4 ]; then
  ab pencil 300 pencilX1024_f64,pencilX512_f64,pencilY2048_f32 default,nccl_c2,nccl_c4,p2p_c1,p2p_c2,p2p_c4,store_c1,store_c2
fi
ls -la $O
