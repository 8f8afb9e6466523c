#!/bin/bash
# 3D sweep kernel variants (openmoc_b200/lib_hint_*.so built with -D experiment macros) on the 3D C5G7 deck
cp openmoc_b200/libb200moc.so /tmp/lib_keep.so
for f in openmoc_b200/lib_hint_*.so; do
  cp $f openmoc_b200/libb200moc.so
  timeout 300 python tools/sweep3d_probe.py "$@" 2>&1 | tail -1 | sed "s|^|[$f] |"
done
cp /tmp/lib_keep.so openmoc_b200/libb200moc.so
timeout 300 python tools/sweep3d_probe.py "$@" 2>&1 | tail -1 | sed "s|^|[base] |"
