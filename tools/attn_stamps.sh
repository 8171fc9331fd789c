#!/bin/bash
# Builds tools/bin/libtvt_stamps.so: the product library with -DTVT_ATTN_STAMPS (phase stamps in the attention backward).
set -e
cd "$(dirname "$0")/.."
CS="data-efficient-video-transformers_b200/csrc"
mkdir -p tools/bin
srcs=""
for f in core gemm_sm100 layernorm attention_simt attention_sm100 pool loss misc; do srcs="$srcs $CS/$f.cu"; done
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  -DTVT_ATTN_STAMPS -shared -o tools/bin/libtvt_stamps.so $srcs
echo built tools/bin/libtvt_stamps.so
