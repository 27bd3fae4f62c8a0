#!/bin/bash
# 8-GPU validation: parity worker (default transports), then the multi-GPU bench lines
NG=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $NG --master-port 29501 tests/gpu_dist_worker.py > gpurun_out/multi_$NG.log 2>&1
echo "worker nproc=$NG rc=$?"; grep -c GPU_WORKER_OK gpurun_out/multi_$NG.log; grep -i "rel L2\|Error\|warn" gpurun_out/multi_$NG.log | head -8
for w in slab1024_f64 slab1024_f64_32 pencilX1024_f64 slab256_f32; do
  timeout 200 $TR --nproc-per-node $NG --master-port 29601 bench.py --gpus $NG --steps 10 --warmup 3 --workload $w --no-e2e \
      > gpurun_out/bench_${w}_$NG.json 2> gpurun_out/bench_${w}_$NG.err
  echo "== $w n=$NG rc=$?"; python scripts/show_passes.py gpurun_out/bench_${w}_$NG.json; grep -v "OMP_NUM\|^\*\*\*\|^$\|^\[W" gpurun_out/bench_${w}_$NG.err | tail -4
done
B200FFT_TRANSPORT=nccl timeout 200 $TR --nproc-per-node $NG --master-port 29602 bench.py --gpus $NG --steps 10 --warmup 3 --workload slab1024_f64 --no-e2e \
      > gpurun_out/bench_slab1024_f64_${NG}_nccl.json 2> gpurun_out/bench_slab1024_f64_${NG}_nccl.err
echo "== slab1024_f64 nccl n=$NG"; python scripts/show_passes.py gpurun_out/bench_slab1024_f64_${NG}_nccl.json
