#!/bin/bash
mkdir -p gpurun_out
GPRN_PROFILE=1 timeout 200 python bench.py --workload c2 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/u_c2_prof.json 2> gpurun_out/u_c2_prof.err; echo "rc=$?"
grep "gprn profile" gpurun_out/u_c2_prof.err | sort -k6 -n -r | head -20
