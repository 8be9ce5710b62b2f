#!/bin/sh
# Writes the SASS evidence VERDICT r1 (missing 6) asked for: the bulk-copy (TMA engine) instructions in the
# shipped cubin of libb200sa.so -- UBLKCP (cp.async.bulk global->shared with mbarrier completion, the
# double-buffered tile loads of msd_partition_kernel) and UBLKPF (cp.async.bulk.prefetch.L2, the next-tile
# prefetch of msd_local_sort_kernel) -- with the function each sits in and a few lines of context.
# usage: tools/sass_excerpt.sh > profiles/r2_sass_tma_excerpt.txt
SO=${1:-stralg_b200/lib/libb200sa.so}
echo "# cuobjdump -sass $SO | functions that contain UBLKCP / UBLKPF / SYNCS (mbarrier) -- $(date -u +%Y-%m-%dT%H:%MZ)"
echo "# nvcc: $(nvcc --version | tail -2 | head -1)"
cuobjdump -sass "$SO" | awk '
/Function :/ { fn=$0 }
/UBLKCP|UBLKPF|SYNCS\.ARRIVE\.TRANS64|SYNCS\.PHASECHK/ { if (fn != last) { print ""; print fn; last=fn } ; print $0 }'
echo
echo "# counts over the whole library"
for op in UBLKCP UBLKPF SYNCS.ARRIVE.TRANS64 SYNCS.PHASECHK; do
  printf "%s: " $op; cuobjdump -sass "$SO" | grep -c "$op"
done
