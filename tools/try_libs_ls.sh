#!/bin/bash
cp openmoc_b200/libb200moc.so /tmp/lib_keep.so
for f in openmoc_b200/lib_hint_*.so; do
  cp $f openmoc_b200/libb200moc.so
  echo "[$f]"; tools/ls_bench.sh
done
cp /tmp/lib_keep.so openmoc_b200/libb200moc.so
echo "[base]"; tools/ls_bench.sh
