#!/bin/bash
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "match" -x 2>&1 | tail -30
