#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -x -k "itransformer" > gpurun_out/pytest_itr_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_itr_$tag.log
grep -E "itransformer_|passed|failed|Error|error|assert" gpurun_out/pytest_itr_$tag.log | tail -30
