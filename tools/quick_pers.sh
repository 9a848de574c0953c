#!/bin/bash
# development aid: 2048-bit shared-operator recurrence, whole-wave launches vs work units
for b in 16384 23680 40000; do
  for f in 0 1; do
    echo "batch $b CB200_PERSISTENT=$f"; CB200_PERSISTENT=$f python tools/quick_time.py 64 $b 200 | tail -n 2 | head -n 1
  done
done
python tools/quick_time.py 128 9472 60 | tail -n 2 | head -n 1
