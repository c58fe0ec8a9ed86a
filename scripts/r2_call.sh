#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -k "decimate or ndvi or golden or scene" > gpurun_out/c_tests.log 2>&1; echo tests $?; tail -2 gpurun_out/c_tests.log
