#!/bin/bash
# CMFD eigenvalue kernel variants (openmoc_b200/lib_hint_cmfd*.so built with -DCMFD_GRID_MIN_BLOCKS=n)
cp openmoc_b200/libb200moc.so /tmp/lib_keep.so
for f in openmoc_b200/lib_hint_cmfd*.so; do
  cp $f openmoc_b200/libb200moc.so
  for t in 128 256; do
    B200_CMFD_MODE=1 B200_CMFD_THREADS=$t timeout 300 python tools/cmfd_bench.py "$@" 2>&1 | tail -1 | sed "s|^|[$f] |"
  done
done
cp /tmp/lib_keep.so openmoc_b200/libb200moc.so
