#!/bin/bash
# end-of-round validation of the final tree: full GPU test suite, then smoke()
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q --durations=4 > gpurun_out/end_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/end_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/end_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/end_smoke.log
