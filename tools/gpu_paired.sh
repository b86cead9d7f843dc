#!/bin/bash
# first GPU pass over the half-matrix path: kernel-level tests, then the headline parity tests, then a short bench
mkdir -p gpurun_out
TAG=${1:-r03a}
timeout 600 python -m pytest tests/test_gpu_paired.py -x -q -s -m gpu > gpurun_out/${TAG}_paired.log 2>&1
echo "paired rc=$?" >> gpurun_out/${TAG}_paired.log
tail -30 gpurun_out/${TAG}_paired.log
