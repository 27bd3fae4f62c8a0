#!/bin/bash
# Round-2 opener, N GPUs (gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_round2_multi.sh N'):
# parity of every transport x pipeline, then the slab 1024^3 bench for each combination.
N=${1:-2}
O=gpurun_out/r02_multi_$N
mkdir -p $O
B200FFT_EXPERIMENTAL=1 timeout 2400 python -m pytest tests/test_gpu_multi.py tests/test_zz_gpu_transports.py -m gpu -q -rxXs \
    -k "test_multi_gpu_parity[$N] or test_slab_transport_parity[$N-" > $O/pytest.log 2>&1
tail -25 $O/pytest.log
# parity of the copy-engine transport with one copy stream per peer (slab, pencil, line)
B200FFT_COPY_STREAMS=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29500 + RANDOM % 500)) tests/gpu_dist_worker.py --transport p2p x > $O/parity_copy_streams.log 2>&1
echo "copy-streams parity rc=$? ($(grep -c GPU_WORKER_OK $O/parity_copy_streams.log) of $N ranks ok)"
run() {  # transport pipeline chunks workload
  B200FFT_TRANSPORT=$1 B200FFT_PIPELINE=$2 B200FFT_CHUNKS=$3 timeout 300 python -m torch.distributed.run --nnodes=1 \
      --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --steps 10 --warmup 3 \
      --no-e2e --no-cpu-baseline --workload $4 > $O/bench_$4_$1_$2_c$3.json 2> $O/bench_$4_$1_$2_c$3.err
  echo "== $4 $1 $2 chunks=$3"; python scripts/show_passes.py $O/bench_$4_$1_$2_c$3.json | tail -4; tail -1 $O/bench_$4_$1_$2_c$3.err
}
for w in slab1024_f64 slab1024_f64_32; do
  run nccl x 0 $w; run p2p x 0 $w; run store x 1 $w
  for c in 2 4 8; do run nccl kz $c $w; run p2p kz $c $w; run store kz $c $w; done
done
# copy-engine transport with one copy stream per peer
for c in 0 4 8; do
  B200FFT_COPY_STREAMS=1 run p2p x $c slab1024_f64
  mv $O/bench_slab1024_f64_p2p_x_c$c.json $O/bench_slab1024_f64_p2p_x_c${c}_cs.json
done
# fused z+y kernel through L2 inside each exchange chunk (groups of 4 / 8 planes), best transports
for g in 4 8; do
  for t in p2p store; do
    B200FFT_L2_MODE=3 B200FFT_L2_PLANES=$g run $t x 0 slab1024_f64
    mv $O/bench_slab1024_f64_${t}_x_c0.json $O/bench_slab1024_f64_${t}_x_c0_l2f$g.json
  done
done
run p2p x 0 slab1024_f64   # (restores the plain result file name)
if [ "$N" -ge 4 ]; then
  for t in nccl p2p store; do run $t x 0 pencilX1024_f64; done
  for c in 2 4; do run nccl x $c pencilX1024_f64; run p2p x $c pencilX1024_f64; done   # pipelined pencil programs
fi
ls -la $O
