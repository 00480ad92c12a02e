#!/bin/bash
# timing experiments for the attention backward kernel (results are WRONG with dbg != 0; timing only)
for d in 0 1 2 4 8 15; do
  echo "dbg=$d"; OCT_ATTN_BWD_DBG=$d timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_bwd_tc_kernel -s 2 -c 1 python tools/profile_attn.py dec 2>&1 | grep -E "duration"
done
