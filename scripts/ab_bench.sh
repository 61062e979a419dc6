#!/bin/bash
# A/B of two builds of the library on the same box, alternating (the boxes are power-capped and differ by a few percent):
#   scripts/ab_bench.sh <libA.so> <libB.so> <out.log> [rounds]
A=$1; B=$2; OUT=$3; N=${4:-3}
for i in $(seq $N); do
  for v in A B; do
    lib=$A; [ $v = B ] && lib=$B
    NBM_B200_LIB=$lib timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-detect --no-stress --parity-clips 1 --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; o=r['other_kernels_ms_per_launch']
print('$v', 'step %.2f ms  slide %.2f  anchor %.2f  tile %.2f  clock %s' % (d['ms_per_step'], r['ms_per_launch'], o['anchor_tc_kernel'], o['tile_kernel'], d['clocks']['sm_mhz']))" >> $OUT
  done
done
