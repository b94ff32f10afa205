#!/bin/bash
# development aid: the tests that exercise the fused / peer-memory training step, then the cfg3 stage table
set -x
timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "two_ranks" 2>&1 | tail -15
timeout 300 python tools/train_profile_cfg3.py 10 adam
timeout 300 python tools/train_profile_cfg3.py 10 sgd
