#!/bin/bash
# Round-2 evidence capture (run under gpurun, 1 GPU):  bash profiles/tools/r2_capture.sh
# Writes ncu reports / launch lists / the clock trace into gpurun_out/; summarise with profiles/summarize_ncu.py.
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
# launch list of the default bench command (kernel share of a step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_launches_c2.log 2>&1
# dominant kernels, one capture each (skip the warm-up launches)
$NCU -k regex:snsde_tc_kernel -s 4 -c 1 -o $O/r2_c2_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > /dev/null 2>&1
$NCU -k regex:snsde_tcg_kernel -s 4 -c 1 -o $O/r2_c4_tcg python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:snsde_tcg_kernel -s 4 -c 1 -o $O/r2_c5_tcg python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
SNSDE_NO_WARP=1 $NCU -k regex:snsde_fma_kernel -s 4 -c 1 -o $O/r2_c1_fma python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:snsde_warp_kernel -s 4 -c 1 -o $O/r2_c1_warp python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline --no-configs > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2_launches_c1.csv python bench.py --workload c1 --no-configs --no-cpu-baseline --steps 5 --warmup 3 > /dev/null 2>&1
$NCU -k regex:snsde_bwd_kernel -s 1 -c 1 -o $O/r2_c2_bwd python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# clock64 trace of the resident kernel (trace build)
SNSDE_TRACE_BUILD=1 SNSDE_TC_TRACE=$O/r2_trace_c2.txt python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > /dev/null 2>&1
python profiles/trace_tc.py $O/r2_trace_c2.txt > $O/r2_trace_c2_summary.txt
