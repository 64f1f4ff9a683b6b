#!/bin/bash
# round-2 GPU pass D: 16 sets in flight, ncu --set full captures (assembly, trsm_col, fused small kernel), DRAM traffic
mkdir -p gpurun_out
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu --pool-per-gpu 16 > gpurun_out/d_c4_p16.json 2> gpurun_out/d_c4_p16.err; echo "c4 p16 rc=$?"; tail -c 900 gpurun_out/d_c4_p16.json; tail -3 gpurun_out/d_c4_p16.err
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu --pool-per-gpu 12 > gpurun_out/d_c4_p12.json 2> gpurun_out/d_c4_p12.err; echo "c4 p12 rc=$?"; tail -c 500 gpurun_out/d_c4_p12.json
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:kassemble_sym -c 1 -f -o gpurun_out/ncu_kassemble_sym python tools/probe_elbo.py 4096 4 2 M52 2 1 > gpurun_out/d_ncu1.log 2>&1; echo "ncu sym rc=$?"
timeout 300 $NCU -k regex:kassemble_rect -c 1 -f -o gpurun_out/ncu_kassemble_rect python tools/probe_predict.py 2048 20000 > gpurun_out/d_ncu2.log 2>&1; echo "ncu rect rc=$?"
timeout 400 $NCU -k regex:trsm_col -s 200 -c 1 -f -o gpurun_out/ncu_trsm_col python tools/probe_elbo.py 4096 4 2 M52 4 1 > gpurun_out/d_ncu3.log 2>&1; echo "ncu trsm rc=$?"
timeout 300 $NCU -k regex:small_pipeline -s 3 -c 1 -f -o gpurun_out/ncu_small_pipeline python tools/probe_elbo.py 256 4 1 QP 4096 3 > gpurun_out/d_ncu4.log 2>&1; echo "ncu small rc=$?"
for it in 1 2; do
  GPRN_NO_GRAPH=1 GPRN_PROBE_NOWARM=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_traffic_it$it.csv python tools/probe_elbo.py 4096 4 2 M52 1 $it > gpurun_out/d_traffic_it$it.log 2>&1; echo "traffic it$it rc=$?"
done
ls -la gpurun_out/*.ncu-rep
