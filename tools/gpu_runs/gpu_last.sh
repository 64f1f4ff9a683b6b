#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/last_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q -k "c3_full or latency_path or continuous or mcmc or elbo_matches_reference or stub_on" > gpurun_out/last_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/last_pytest.log
