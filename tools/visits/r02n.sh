#!/bin/bash
SFB_ATTN_TIMING=1 timeout 120 python tools/gpu_microbench.py attn_self_S32760 attn_self_S4680 2>&1 | grep "attn timing" | sort | uniq -c | sort -rn | head -6
