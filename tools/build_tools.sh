#!/bin/bash
# Builds the stand-alone kernel bring-up harnesses into tools/bin/ (git-ignored, shipped by gpurun).
set -e
cd "$(dirname "$0")/.."
CS="data-efficient-video-transformers_b200/csrc"
mkdir -p tools/bin
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17"
$NV -o tools/bin/gemm_check tools/gemm_check.cu $CS/gemm_sm100.cu $CS/core.cu
echo built tools/bin/gemm_check
