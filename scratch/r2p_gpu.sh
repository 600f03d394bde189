#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_colocated.py -x -q -m gpu -k "linear or tower or deepfm or DeepFM or group or colocat" 2>&1 | tail -4
bash scratch/ab_bench.sh RLCTR_MLP_REUSE_SPLIT 0 1 2>&1 | tail -4
