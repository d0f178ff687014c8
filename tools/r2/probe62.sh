#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/probe62_tests.log 2>&1
tail -5 gpurun_out/probe62_tests.log
