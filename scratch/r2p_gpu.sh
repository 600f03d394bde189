#!/bin/bash
# round-2 late: GEMM changes (k-step trimming, epilogue-warp B split in wgrad, vectorized single-output layer)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear or tower or mlp" > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?" 
tail -3 gpurun_out/r2p_tests.log
export GEMM_BENCH_FAST=1
echo "== default"; timeout 120 python scratch/gemm_bench.py 2>&1 | tail -2
echo "== BSPLIT_EPI=0"; RLCTR_GEMM_BSPLIT_EPI=0 timeout 120 python scratch/gemm_bench.py 2>&1 | tail -2
echo "== NT_MAX=128"; RLCTR_GEMM_NT_MAX=128 timeout 120 python scratch/gemm_bench.py 2>&1 | tail -2
echo "== CLUSTER=2"; RLCTR_GEMM_CLUSTER=2 timeout 120 python scratch/gemm_bench.py 2>&1 | tail -2
