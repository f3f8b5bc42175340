#!/bin/bash
# ncu evidence for K1 alone (default workload): per-colour DRAM bytes / FP64 pipe, one full capture with the source page.
# Usage under gpurun: bash tools/profile_k1.sh <tag>   -> gpurun_out/<tag>_k1_{traffic,raw,source}.csv
set -u
T=${1:-r02b}
O=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-parity"
NCU="ncu --clock-control none"
$B > $O/${T}_plain.log 2>&1 || { echo "plain bench failed"; tail -5 $O/${T}_plain.log; exit 1; }
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum \
     -k regex:k_assemble_regular -c 7 --csv --log-file $O/${T}_k1_traffic.csv $B > $O/${T}_ncu2.log 2>&1; echo "K1 traffic rc=$?"
$NCU --set full --import-source on -k regex:k_assemble_regular -s 1 -c 1 -o $O/${T}_k1 -f $B > $O/${T}_ncu3.log 2>&1; echo "K1 full rc=$?"
ncu -i $O/${T}_k1.ncu-rep --page raw --csv > $O/${T}_k1_raw.csv 2>/dev/null
ncu -i $O/${T}_k1.ncu-rep --page source --csv > $O/${T}_k1_source.csv 2>/dev/null
rm -f $O/${T}_k1.ncu-rep
ls -la $O/${T}_*
