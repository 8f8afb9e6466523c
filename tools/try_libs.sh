#!/bin/bash
cp openmoc_b200/libb200moc.so /tmp/lib_keep.so
for f in openmoc_b200/lib_hint_*.so; do
  cp $f openmoc_b200/libb200moc.so
  timeout 120 python tools/sweep_tune.py "$@" 2>&1 | tail -1 | sed "s|^|[$f] |"
done
cp /tmp/lib_keep.so openmoc_b200/libb200moc.so
timeout 120 python tools/sweep_tune.py "$@" 2>&1 | tail -1 | sed "s|^|[base] |"
