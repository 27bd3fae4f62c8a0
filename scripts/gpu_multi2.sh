#!/bin/bash
# multi-GPU: parity worker with forced chunking, then bench vs pipeline depth
NG=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B200FFT_CHUNKS=2 timeout 600 $TR --nproc-per-node $NG --master-port 29501 tests/gpu_dist_worker.py > gpurun_out/multi_c2_$NG.log 2>&1
echo "worker chunks=2 nproc=$NG rc=$?"; grep -c GPU_WORKER_OK gpurun_out/multi_c2_$NG.log; grep "rel L2" gpurun_out/multi_c2_$NG.log | head -5
timeout 600 $TR --nproc-per-node $NG --master-port 29502 tests/gpu_dist_worker.py > gpurun_out/multi_auto_$NG.log 2>&1
echo "worker auto nproc=$NG rc=$?"; grep -c GPU_WORKER_OK gpurun_out/multi_auto_$NG.log; grep "rel L2" gpurun_out/multi_auto_$NG.log | head -5
for w in slab1024_f64 slab1024_f64_32; do
  for c in 1 2 4 8; do
    B200FFT_CHUNKS=$c timeout 300 $TR --nproc-per-node $NG --master-port $((29600+c)) bench.py --gpus $NG --steps 10 --warmup 3 --workload $w --no-e2e \
      > gpurun_out/bench_${w}_${NG}_c$c.json 2> gpurun_out/bench_${w}_${NG}_c$c.err
    echo "== $w n=$NG chunks=$c rc=$?"; python scripts/show_passes.py gpurun_out/bench_${w}_${NG}_c$c.json; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_${w}_${NG}_c$c.err | tail -3
  done
done
