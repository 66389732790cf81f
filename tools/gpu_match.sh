#!/bin/bash
timeout 300 python tools/time_match.py && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_match.csv -k regex:match python tools/time_match.py > /dev/null 2>&1
timeout 300 python -m pytest tests -m gpu -q -k match -p no:cacheprovider 2>&1 | tail -3
