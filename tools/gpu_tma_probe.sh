#!/bin/bash
# TMA patch-load probe: bw K step nblk nph align(2 = 16-aligned interior, 1 = 4-aligned, 0 = any, may leave the image) swz
for cfg in "16 40 1 1 1 2 0" "16 40 1 5 2 2 0" "16 40 1 5 2 2 1" "16 40 2 5 2 2 1" "16 40 2 5 2 1 1" "16 40 2 5 2 0 1" "16 48 1 4 1 0 1" "32 40 2 3 2 0 1" "16 40 3 5 2 0 1" "16 32 2 4 2 0 1" "32 32 2 2 2 0 1" "16 36 4 5 2 0 1"; do
  timeout 100 build/tma_probe $cfg 2>&1 | tail -8
done
