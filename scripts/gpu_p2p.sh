#!/bin/bash
# copy-engine transport: parity (forced chunking too) and bench vs NCCL at NG GPUs
NG=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for c in 0 2; do
  B200FFT_TRANSPORT=p2p B200FFT_CHUNKS=$c timeout 300 $TR --nproc-per-node $NG --master-port $((29510+c)) tests/gpu_dist_worker.py > gpurun_out/p2p_worker_c$c.log 2>&1
  echo "p2p worker chunks=$c nproc=$NG rc=$?"; grep -c GPU_WORKER_OK gpurun_out/p2p_worker_c$c.log; grep -i "rel L2\|error\|Traceback" gpurun_out/p2p_worker_c$c.log | head -8
done
for cfg in "nccl 1" "nccl 2" "p2p 1" "p2p 2" "p2p 4" "p2p 8"; do
  set -- $cfg
  B200FFT_TRANSPORT=$1 B200FFT_CHUNKS=$2 timeout 200 $TR --nproc-per-node $NG --master-port $((29600+$2)) bench.py --gpus $NG --steps 10 --warmup 3 --workload slab1024_f64 --no-e2e \
      > gpurun_out/bench_${NG}_$1_c$2.json 2> gpurun_out/bench_${NG}_$1_c$2.err
  echo "== slab1024_f64 n=$NG transport=$1 chunks=$2 rc=$?"; python scripts/show_passes.py gpurun_out/bench_${NG}_$1_c$2.json; grep -v "OMP_NUM\|^\*\*\*\|^$\|^\[W" gpurun_out/bench_${NG}_$1_c$2.err | tail -4
done
