#!/bin/bash
# Development aid (runs here, no GPU): turn an .ncu-rep brought back in gpurun_out/ into the text
# summaries committed under profiles/ -- the details page and the per-source-line table
# (tools/ncu_lines.py) of every round-0 kernel in the report.
#   tools/summarise_ncu.sh gpurun_out/r1_msd_final.ncu-rep profiles/r1
set -e
REP=$1; OUT=$2
for K in msd_local_sort_kernel msd_partition_text_kernel msd_partition_kernel; do
  S=$K
  [ $K = msd_partition_kernel ] && S="msd_partition_kernelENS"
  ncu -i $REP --page details --kernel-name regex:"^${K}\$|^${K}\(|${K}<" 2>/dev/null | grep -v '^ *$' | cut -c1-170 > ${OUT}_${K}_details.txt || true
  NCU_NAME="^${K}" python tools/ncu_lines.py $REP $S 40 > ${OUT}_${K}_lines.txt 2>&1 || true
done
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; un=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__grid_size','launch__block_size']
print(','.join(want))
for d in rows[2:]:
    print(','.join('\"'+d[hdr.index(w)]+'\"' if w in hdr else '' for w in want))
" > ${OUT}_summary.csv
ls -la ${OUT}_*
