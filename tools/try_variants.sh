#!/bin/bash
# usage: tools/try_variants.sh  (on the GPU box)
for v in regs ring; do
  B200_SWEEP=$v timeout 120 python tools/sweep_tune.py 2>&1 | tail -1 | sed "s/^/[$v] /"
  B200_SWEEP=$v timeout 120 python tools/sweep_tune.py --precision mixed 2>&1 | tail -1 | sed "s/^/[$v] /"
done
cp openmoc_b200/libb200moc.so /tmp/lib_keep.so; cp openmoc_b200/libb200moc_minb3.so openmoc_b200/libb200moc.so
B200_SWEEP=ring timeout 120 python tools/sweep_tune.py 2>&1 | tail -1 | sed "s/^/[ring minb3] /"
B200_SWEEP=ring timeout 120 python tools/sweep_tune.py --precision mixed 2>&1 | tail -1 | sed "s/^/[ring minb3] /"
cp /tmp/lib_keep.so openmoc_b200/libb200moc.so
