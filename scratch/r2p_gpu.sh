#!/bin/bash
mkdir -p gpurun_out
export GEMM_BENCH_FAST=1
for m in 0 1 2 0 1 2; do echo "== MMA_SOLO=$m"; RLCTR_GEMM_MMA_SOLO=$m timeout 120 python scratch/gemm_bench.py 2>&1 | tail -2 | cut -c1-330; done
