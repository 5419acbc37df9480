#!/bin/bash
# One GPU call: smoke, the -m gpu suite, the default bench line and the reference arm (run through gpurun).
set -x
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
cat gpurun_out/c1_parity.json
