#!/bin/bash
timeout 25 python -m pytest tests/test_gpu_kiez.py -m gpu -q -k "sort_mirrors or api_behaviour" -p no:cacheprovider 2>&1 | tail -3
