#!/bin/bash
# One gpurun --gpus N call: multi-GPU parity worker at 2 and N ranks, then multi-rank bench lines.
NG=${1:-4}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 $NG; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29500+n)) tests/gpu_dist_worker.py > gpurun_out/multi_$n.log 2>&1
  echo "worker nproc=$n rc=$?" | tee -a gpurun_out/multi_$n.log
  tail -4 gpurun_out/multi_$n.log
done
for n in 2 $NG; do
  for w in slab1024_f64 slab1024_f64_32; do
    timeout 300 $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 --workload $w --no-e2e \
      > gpurun_out/bench_${w}_$n.json 2> gpurun_out/bench_${w}_$n.err
    echo "bench $w n=$n rc=$?"; tail -c 1500 gpurun_out/bench_${w}_$n.json; tail -3 gpurun_out/bench_${w}_$n.err
  done
done
if [ $NG -ge 4 ]; then
  timeout 300 $TR --nproc-per-node $NG --master-port 29700 bench.py --gpus $NG --steps 10 --warmup 3 --workload pencilX1024_f64 --no-e2e \
      > gpurun_out/bench_pencilX1024_$NG.json 2> gpurun_out/bench_pencilX1024_$NG.err
  echo "bench pencil rc=$?"; tail -c 1500 gpurun_out/bench_pencilX1024_$NG.json; tail -3 gpurun_out/bench_pencilX1024_$NG.err
fi
