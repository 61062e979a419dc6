#!/bin/bash
# the same for N builds: scripts/abn_bench.sh out.log rounds lib1.so lib2.so ...
OUT=$1; N=$2; shift 2
for i in $(seq $N); do
  for lib in "$@"; do
    NBM_B200_LIB=$lib timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-detect --no-stress --parity-clips 1 --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; o=r['other_kernels_ms_per_launch']
print('$(basename $lib)', 'step %.2f ms  slide %.2f  anchor %.2f  tile %.2f  clock %s  parity %.2e' % (d['ms_per_step'], r['ms_per_launch'], o['anchor_tc_kernel'], o['tile_kernel'], d['clocks']['sm_mhz'], d['parity']['max_abs_err_norm']))" >> $OUT
  done
done
