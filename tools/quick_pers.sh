#!/bin/bash
# development aid: 2048-bit shared-operator recurrence, work-unit kernel with 1 or 2 CTAs per SM
for b in 14208 16384 28416 40000; do
  for c in 1 2 3; do
    echo "batch $b CB200_PERS_CTAS=$c"; CB200_PERS_CTAS=$c python tools/quick_time.py 64 $b 200 | tail -n 2 | head -n 1
  done
done
