#!/bin/bash
# Runs the conv bring-up harness on the GPU box; every case in its own process under a timeout.
mkdir -p gpurun_out
LOG=gpurun_out/selftest.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
B=build/conv_selftest
echo "== diag" >> $LOG
timeout 60 $B diag >> $LOG 2>&1; echo "exit=$?" >> $LOG
read NC NP NWC NWP < <($B list)
for i in $(seq 0 $((NC-1))); do
  echo "== case $i" >> $LOG
  timeout 60 $B case $i >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
for i in $(seq 0 $((NP-1))); do
  echo "== perf $i" >> $LOG
  timeout 120 $B perf $i >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
for i in $(seq 0 $((NWC-1))); do
  echo "== wcase $i" >> $LOG
  timeout 60 $B wcase $i >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
for i in $(seq 0 $((NWP-1))); do
  echo "== wperf $i" >> $LOG
  timeout 120 $B wperf $i >> $LOG 2>&1; echo "exit=$?" >> $LOG
done
grep -E "PASS|FAIL|exit=[1-9]|TFLOP" $LOG | tail -120
