#!/bin/bash
for c in 16 32 74 148 296 592; do echo "ctas=$c"; SB_DESC_CTAS=$c timeout 300 python tools/prof_kernels.py 8 1 | cut -c1-200; SB_DESC_CTAS=$c timeout 300 python tools/prof_kernels.py 64 1 | cut -c1-200; done
