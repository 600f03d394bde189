#!/bin/bash
mkdir -p gpurun_out
export GEMM_BENCH_FAST=1
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear or tower" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "tower or deepfm or DeepFM" 2>&1 | tail -2
for i in 1 2; do timeout 120 python scratch/gemm_bench.py 2>&1 | tail -2 | cut -c1-420; done
