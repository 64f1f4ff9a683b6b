#!/bin/bash
mkdir -p gpurun_out
{
python tools/latency_probe.py 500 4 40
GPRN_NO_LOOP=1 python tools/latency_probe.py 500 4 40
GPRN_NO_LOOP=1 GPRN_MID_WARPS=4 python tools/latency_probe.py 500 4 40
GPRN_NO_MID=1 python tools/latency_probe.py 500 4 10
python tools/latency_probe.py 130 1 40
GPRN_NO_LOOP=1 python tools/latency_probe.py 130 1 40
} > gpurun_out/y_latency.txt 2>&1
cat gpurun_out/y_latency.txt
