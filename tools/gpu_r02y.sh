#!/bin/bash
# round 2, GPU call Y (1 GPU): the early-fetch test on the final library (the Finder tests and smoke() ran green on it in the
# first attempt of this call, whose early-fetch test compared group ORDER between two builds -- which is not defined)
set -u
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_build.py -q -m gpu -k "early_fetch" > gpurun_out/r02y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02y_pytest.log
