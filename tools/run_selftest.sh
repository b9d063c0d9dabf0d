#!/bin/bash
# usage: tools/run_selftest.sh group1 group2 ...   (each group in its own process, 300 s cap)
mkdir -p gpurun_out
for g in "$@"; do
  echo "=== $g ===" | tee -a gpurun_out/selftest.log
  timeout 60 python tools/selftest.py "$g" 2>&1 | tail -80 | tee -a gpurun_out/selftest.log
  echo "exit=${PIPESTATUS[0]}" | tee -a gpurun_out/selftest.log
done
