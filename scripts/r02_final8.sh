#!/bin/bash
# Round 2, final 8-GPU run: default configuration of every BASELINE workload (forward parity inside), flag kernel
# against per-peer DMA flags, the official bench line of the headline workload.
N=8
O=gpurun_out/r02_final_8
mkdir -p $O
port() { echo $((29500 + RANDOM % 2000)); }
trun() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) "${@:2}"; }
trun 150 scripts/ab_multi.py --steps 20 --workloads slab1024_f64,slab1024_f64_32,pencilX512_f64,pencilX1024_f64,pencilY2048_f32,line16384_f32,slab256_f32 \
    --configs default,p2p_kz2 > $O/ab_defaults.jsonl 2> $O/ab_defaults.txt
echo "ab defaults rc=$?" | tee -a $O/summary.txt; grep "^slab\|^pencil\|^line" $O/ab_defaults.txt | sed 's/^/flag kernel: /'
B200FFT_FLAG_DMA=1 trun 100 scripts/ab_multi.py --steps 20 --workloads slab1024_f64,pencilX1024_f64 --configs default,p2p_kz2 > $O/ab_flag_dma.jsonl 2> $O/ab_flag_dma.txt
echo "ab flag dma rc=$?" | tee -a $O/summary.txt; grep "^slab\|^pencil\|^line" $O/ab_flag_dma.txt | sed 's/^/flag DMA:    /'
trun 100 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_slab1024_f64.json 2> $O/bench.err
python scripts/show_passes.py $O/bench_slab1024_f64.json
