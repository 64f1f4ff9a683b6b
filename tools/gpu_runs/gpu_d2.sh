#!/bin/bash
mkdir -p gpurun_out
GPRN_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:mid_pipeline -s 12 -c 2 -o gpurun_out/d2_mid_full -f python tools/latency_probe.py 500 4 2 > gpurun_out/d2_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -5 gpurun_out/d2_ncu_full.log; ls -la gpurun_out/d2_mid_full.ncu-rep
