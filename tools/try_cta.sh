#!/bin/bash
# CTA size x register cap variants of the flat sweep kernel (lib_hint_t<threads>b<blocks>.so built with
# -DB200_LB_THREADS/-DB200_LB_BLOCKS); B200_CTA sets the launch block size at run time
cp openmoc_b200/libb200moc.so /tmp/lib_keep.so
for f in openmoc_b200/lib_hint_t*.so; do
  t=$(echo $f | sed 's/.*lib_hint_t\([0-9]*\)b.*/\1/')
  cp $f openmoc_b200/libb200moc.so
  B200_CTA=$t timeout 120 python tools/sweep_tune.py "$@" 2>&1 | tail -1 | sed "s|^|[$f CTA=$t] |"
done
cp /tmp/lib_keep.so openmoc_b200/libb200moc.so
for t in 224 192 160 128 96; do
  B200_CTA=$t timeout 120 python tools/sweep_tune.py "$@" 2>&1 | tail -1 | sed "s|^|[base(224,4) CTA=$t] |"
done
