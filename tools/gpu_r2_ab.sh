#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in 4 8 16 32; do
  XEMO_POOLBWD_WAVES=$w timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown_pw$w.txt 2>&1
  echo "== pool bwd waves $w"; grep -- "---- total" gpurun_out/op_breakdown_pw$w.txt
  awk '/---- total/{f=1} f' gpurun_out/op_breakdown_pw$w.txt | grep "maxpool_bwd"
done
