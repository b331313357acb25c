#!/bin/bash
# bring-up of the CTA-pair GEMM on the GPU box: correctness per tile width, then perf v1 vs v2
cd "$(dirname "$0")/../.."
for bn in 128 192 256; do
  echo "=== v2 correctness BN=$bn"
  UVC_GEMM_V2=2 UVC_GEMM_V2_BN=$bn timeout 300 python tests/bringup/gemm_cases.py v2 2>&1 | tail -25
done
echo "=== perf v1"; UVC_GEMM_V2=0 timeout 300 python tests/bringup/gemm_cases.py perf2 2>&1 | grep perf2
for bn in 128 192 256; do
  echo "=== perf v2 BN=$bn"; UVC_GEMM_V2=2 UVC_GEMM_V2_BN=$bn timeout 300 python tests/bringup/gemm_cases.py perf2 2>&1 | grep perf2
done
echo "=== perf v2 auto"; UVC_GEMM_V2=1 timeout 300 python tests/bringup/gemm_cases.py perf2 2>&1 | grep perf2
