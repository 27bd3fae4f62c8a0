#!/bin/bash
# 4 GPUs: pipeline depth / direction sweep of the slab exchange (the default picked 8 chunks here without a measurement)
N=4
O=gpurun_out/r02_sweep_4
mkdir -p $O
port() { echo $((29500 + RANDOM % 2000)); }
trun() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) "${@:2}"; }
trun 150 scripts/ab_multi.py --steps 20 --workloads slab1024_f64,slab1024_f64_32 --configs default,p2p_c2,p2p_c4,p2p_c8,p2p_kz2,p2p_kz4 > $O/ab.jsonl 2> $O/ab.txt
grep "^slab" $O/ab.txt
