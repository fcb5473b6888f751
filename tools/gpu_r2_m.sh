#!/bin/bash
# CTA-pair (cta_group::2) bring-up: every correctness case of the harness with pairs forced, then the timing cases both ways.
mkdir -p gpurun_out
B=build/conv_selftest
read NC NP NWC NWP < <($B list)
LOG=gpurun_out/selftest_2cta.log; : > $LOG
for i in $(seq 0 $((NC-1))); do
  echo "== case $i (pairs forced)" >> $LOG
  XEMO_CONV_2CTA=2 timeout 60 $B case $i >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
for i in $(seq 0 $((NP-1))); do
  for m in 0 2; do
    echo "== perf $i XEMO_CONV_2CTA=$m" >> $LOG
    XEMO_CONV_2CTA=$m timeout 120 $B perf $i >> $LOG 2>&1; echo "exit=$?" >> $LOG
  done
done
grep -E "==|PASS|FAIL|exit=[1-9]|TFLOP|error|rror" $LOG | tail -150
