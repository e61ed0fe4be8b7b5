#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged_geometry" 2>&1 | tail -15
