#!/bin/bash
mkdir -p gpurun_out
{
python tools/latency_probe.py 500 4 40
GPRN_B200_LIB=$PWD/gpurun_in/libgprn_nopf.so python tools/latency_probe.py 500 4 40
python tools/latency_probe.py 1000 4 10
GPRN_B200_LIB=$PWD/gpurun_in/libgprn_nopf.so python tools/latency_probe.py 1000 4 10
GPRN_MID_WARPS=4 python tools/latency_probe.py 500 4 40
} > gpurun_out/f2_latency.txt 2>&1
cat gpurun_out/f2_latency.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "c3_full or mid_n or up_to_n_1024" > gpurun_out/f2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f2_pytest.log
