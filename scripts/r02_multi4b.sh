#!/bin/bash
# 4 GPUs: the parity worker's transport / pipeline modes (tests/test_zz_gpu_transports.py), one torchrun each
N=4
O=gpurun_out/r02_multi_4
mkdir -p $O
port() { echo $((29500 + RANDOM % 2000)); }
trun() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) "${@:2}"; }
for mode in ${MODES:-store_x p2p_kz store_kz nccl_kz nccl_pencil-chunks p2p_pencil-chunks}; do
  trun 120 tests/gpu_dist_worker.py --transport ${mode%%_*} ${mode#*_} > $O/parity_$mode.log 2>&1
  echo "worker --transport ${mode%%_*} ${mode#*_} rc=$? ($(grep -c GPU_WORKER_OK $O/parity_$mode.log) of $N ranks ok)" | tee -a $O/summary_modes.txt
  grep "^\[" $O/parity_$mode.log | tail -1
done
