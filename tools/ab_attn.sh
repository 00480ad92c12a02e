#!/bin/bash
# times the decoder-shape attention kernels with every experiment build in octcubem_b200/variants (CUDA events, not ncu)
cd "$(dirname "$0")/.."
echo "== default"; python tools/profile_attn.py dec time | grep TIME
for so in octcubem_b200/variants/lib*.so; do
  echo "== $so"; OCT_LIB=$PWD/$so timeout 120 python tools/profile_attn.py dec time | grep TIME
done
