#!/bin/bash
mkdir -p gpurun_out
bash scripts/ab_bench.sh "TREEDET_TRACE_MINBLOCKS=24|" "TREEDET_TRACE_MINBLOCKS=28|" "|" "TREEDET_TRACE_MINBLOCKS=24|"
