#!/bin/bash
# Round 2, final one-GPU check, as the driver runs things: GPU suite, smoke(), the default bench line (with e2e and
# cpu_baseline) and the reference arm.
O=gpurun_out/r02_final_1
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu -rxXs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err ) 2>&1 | grep real
python scripts/show_passes.py $O/bench_n1.json; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print({k: d[k] for k in ('forward_rel_l2','gpu_launches','clocks')}); print('e2e', d['e2e']); print('cpu', d['cpu_baseline']); print('roofline', {k: d['roofline'][k] for k in ('kernel','achieved','frac','traffic','sum_fft_ms')})"
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err ) 2>&1 | grep real
cat $O/bench_ref.json | cut -c1-900
