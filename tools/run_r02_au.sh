#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02au_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02au_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02au_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/r02au_smoke.log
