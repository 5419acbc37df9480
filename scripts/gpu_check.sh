#!/bin/bash
# One GPU call: smoke and the -m gpu suite (run through gpurun).
set -x
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25
