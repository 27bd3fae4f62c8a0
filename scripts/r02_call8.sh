#!/bin/bash
# ncu: launch list of the default bench and full-set capture of one round trip (raw CSV page only comes back; the
# .ncu-rep stays on the box: gpurun_out is limited to 64 MiB)
O=gpurun_out/r02_ncu
mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft_ -c 80 --csv --log-file $O/launches_1024.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launch.log 2>&1
for w in slab1024_f64 slab1024_f64_32; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:fft_ -s 21 -c 6 -o /tmp/prof_$w -f \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload $w > $O/ncu_full_$w.log 2>&1
  ncu -i /tmp/prof_$w.ncu-rep --page raw --csv > $O/prof_$w.raw.csv 2>/dev/null
  ncu -i /tmp/prof_$w.ncu-rep --page details --csv > $O/prof_$w.details.csv 2>/dev/null
done
ls -la $O
