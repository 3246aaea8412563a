#!/bin/bash
# Reconstruction-only A/B of library variants on one box: tools/recon_ab.sh lib lib_r5 ...
for round in 1 2; do
  for v in "$@"; do
    echo -n "$v round $round: "
    HIJIKI_B200_LIB=$PWD/hijiki_b200/$v/libhijiki_b200.so python tools/denoise_probe.py random 50 2>&1 | tail -1
  done
done
