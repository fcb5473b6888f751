#!/bin/bash
# cluster size of the SE gate at small batches (teacher forward graph)
cd "$(dirname "$0")/.."
for b in 32 64 128; do
  for k in 1 2 4 8; do
    XEMO_SE_GATE_K=$k timeout 300 python tools/ab_options.py $b gate 2>&1 | grep teacher | sed "s/^/B=$b /"
  done
done
