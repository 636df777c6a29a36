#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_final_tests.log; cat gpurun_out/r2_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
