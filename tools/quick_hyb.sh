#!/bin/bash
# development aid: all-shared against hybrid scratch columns at whole-wave sizes of both kernels
for h in 0 1; do
  echo "CB200_HYBRID=$h"
  CB200_HYBRID=$h python tools/quick_general.py 128 198912 40 | grep "^frobenius" | tail -n 1
  CB200_HYBRID=$h python tools/quick_general.py 64 113664 40 | grep "^frobenius" | tail -n 1
  CB200_HYBRID=$h python tools/quick_general.py 32 66304 40 | grep "^frobenius" | tail -n 1
done
