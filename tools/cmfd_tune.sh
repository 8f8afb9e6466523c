#!/bin/bash
# launch shapes of the CMFD eigenvalue kernel: device time per SOR iteration on a 2D and a 3D C5G7 mesh
D=oracle/_ref/ref_driver
A="--model c5g7-2d --azim 8 --spacing 0.2 --cmfd 51x51 --max-iters 60"
B="--model c5g7-2d --dims 3 --azim 4 --polar 2 --spacing 1.0 --zspacing 5 --axial 9 --formation otf-stacks --cmfd 51x51x9 --max-iters 12 --threads 8"
one() {   # label, env..., -- args
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" $D "$@" --solver b200 --quiet --json /tmp/t.json > /dev/null 2>&1
  python - "$label" <<'PY'
import json,sys
r=json.load(open('/tmp/t.json'))
n=max(r['cmfd_sor_iterations'],1)
print("%-34s iters %3d k %.9f cmfd kernels %.4f s, %6d SOR its, %6.2f us per SOR iteration" % (sys.argv[1], r['iterations'], r['keff'], r['cmfd_kernels_s'], n, r['cmfd_kernels_s']/n*1e6))
PY
}
for deck in A B; do
  args=${!deck}
  echo "== deck $deck: $args"
  one "default" X=1 -- $args
  for c in 4 8 16; do one "cluster C=$c" B200_CMFD_MODE=2 B200_CMFD_CLUSTER=$c -- $args; done
  for t in 32 64 128 256; do one "grid threads=$t" B200_CMFD_MODE=1 B200_CMFD_THREADS=$t -- $args; done
done
