#!/bin/bash
# first GPU contact: parity tests of the warp32 engine
cd "$(dirname "$0")"
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv
python -m pytest tests/test_gpu_forward_parity.py -m gpu -q 2>&1 | tail -30
