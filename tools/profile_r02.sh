#!/bin/bash
# ncu evidence of round 2 (run under gpurun on one B200; every ncu command follows the same command line exiting 0 without ncu).
# Outputs go to gpurun_out/; the csv pages are exported on the CPU box (tools/export_profiles.sh) and summarised in profiles/.
set -u
O=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-parity"
NCU="ncu --clock-control none"
$B > $O/p_plain.log 2>&1 || { echo "plain bench failed"; tail -5 $O/p_plain.log; exit 1; }
# 1. every launch of the default step with its device time
$NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/r02_launches_c4_m76.csv $B > $O/p_ncu1.log 2>&1; echo "launch list rc=$?"
# 2. the 7 colour launches of one assembly: DRAM bytes, duration, FP64 pipe
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum \
     -k regex:k_assemble_regular -c 7 --csv --log-file $O/r02_k_assemble_regular_c4_m76_traffic.csv $B > $O/p_ncu2.log 2>&1; echo "K1 traffic rc=$?"
# 3. full capture of the second colour launch of K1 (source page needs -lineinfo: build.py has it)
$NCU --set full --import-source on -k regex:k_assemble_regular -s 1 -c 1 -o $O/r02_k1 -f $B > $O/p_ncu3.log 2>&1; echo "K1 full rc=$?"
# 4. one GMRES iteration in the middle of the solve: matvec, three Gram-Schmidt passes, publish
$NCU --set full -k "regex:k_gemv|k_gm_pass|k_gm_publish|k_multi_dot|k_sum_partials" -s 300 -c 8 -o $O/r02_gmres_iter -f $B > $O/p_ncu4.log 2>&1; echo "GMRES iteration rc=$?"
# 5. LU: trailing update on the tensor path and the cooperative application
L="python tests/lu_bench.py 18432"
$L > $O/p_lu_plain.log 2>&1 && \
$NCU --set full -k regex:k_lu_gemm_dmma -s 60 -c 1 -o $O/r02_lu_gemm -f $L > $O/p_ncu5.log 2>&1; echo "LU gemm rc=$?"
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_lu_apply_coop -c 2 --csv --log-file $O/r02_k_lu_apply_coop.csv $L > $O/p_ncu6.log 2>&1; echo "LU apply rc=$?"
# 6. image kernels and Q2 (BASELINE configs 5 and 3) at the sizes of bench.py's secondary workloads
for W in c5 c5ns; do
  C="python bench.py --workload $W --subdiv 48 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-parity"
  $C > $O/p_${W}_plain.log 2>&1 && \
  $NCU --set full -k regex:k_assemble_regular -s 14 -c 1 -o $O/r02_k1_$W -f $C > $O/p_ncu_$W.log 2>&1; echo "$W rc=$?"
done
C="python bench.py --workload q2 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-parity"
$C > $O/p_q2_plain.log 2>&1 && \
$NCU --set full -k regex:k_assemble_regular -s 7 -c 1 -o $O/r02_k1_q2 -f $C > $O/p_ncu_q2.log 2>&1; echo "q2 rc=$?"
# export the pages on the box (gpurun merges at most 64 MiB back) and drop the reports
for R in r02_k1 r02_gmres_iter r02_lu_gemm r02_k1_c5 r02_k1_c5ns r02_k1_q2; do
  [ -f $O/$R.ncu-rep ] || continue
  ncu -i $O/$R.ncu-rep --page raw --csv > $O/${R}_raw.csv 2>/dev/null
  rm -f $O/$R.ncu-rep.keep
done
ncu -i $O/r02_k1.ncu-rep --page source --csv > $O/r02_k1_source.csv 2>/dev/null
ncu -i $O/r02_k1_c5.ncu-rep --page source --csv > $O/r02_k1_c5_source.csv 2>/dev/null
rm -f $O/*.ncu-rep
ls -la $O/r02_* 2>/dev/null
