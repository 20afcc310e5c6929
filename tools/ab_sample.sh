#!/bin/bash
# Same-box A/B of DDPM.sample (64 x 50 steps) across builds of the library: tools/ab_sample.sh <lib.so> [<lib.so> ...]; two interleaved rounds.
for round in 1 2; do
  for lib in "$@"; do
    echo -n "$(basename $lib): "
    LDMB_LIB_PATH=$PWD/$lib python tools/time_sample.py 2>&1 | tail -1
  done
done
