#!/bin/bash
# ncu captures of one colour launch of K1 for the image kernels at the size of bench.py's secondary workloads
# (after the same command exited 0 without ncu).  Usage under gpurun: bash tools/profile_c5.sh
set -u
O=gpurun_out
NCU="ncu --clock-control none"
for W in c5 c5ns; do
  C="python bench.py --workload $W --subdiv 48 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-parity"
  $C > $O/p_${W}_plain.log 2>&1 || { echo "$W plain failed"; continue; }
  $NCU --set full --import-source on -k regex:k_assemble_regular -s 12 -c 1 -o $O/r02_k1_$W -f $C > $O/p_ncu_$W.log 2>&1; echo "$W rc=$?"
  ncu -i $O/r02_k1_$W.ncu-rep --page raw --csv > $O/r02_k1_${W}_raw.csv 2>/dev/null
  ncu -i $O/r02_k1_$W.ncu-rep --page source --csv > $O/r02_k1_${W}_source.csv 2>/dev/null
  rm -f $O/r02_k1_$W.ncu-rep
done
ls -la $O/r02_k1_c5*
