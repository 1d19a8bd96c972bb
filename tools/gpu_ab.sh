#!/bin/bash
for lib in tools/ab/lib_a.so tools/ab/lib_b.so ""; do
  echo "== lib ${lib:-current}"
  BMC_B200_LIB=${lib:+$PWD/$lib} timeout 300 python tools/gpu_diag.py convperf 2>&1 | grep -E "taps=9.*B=19 impl=0|taps=9.*B=16 impl=0"
done
