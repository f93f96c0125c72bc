#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fused_decode.py -x -q -m gpu -k "program" 2>&1 | tail -8
