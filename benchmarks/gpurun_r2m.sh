#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2m_pytest.txt
