#!/bin/bash
# time of the gather descriptor kernel per geometry class: the classes in the mask are NOT described
for m in 0 1 2 4 8 16 32 64 127; do echo "mask $m"; SB_CLS_MASK=$m timeout 100 python tools/prof_kernels.py 64 1 | sed 's/.*per frame us//'; done
