#!/bin/bash
# groups per thread (B200_GPL) on the 70-group decks: 3D lattice (C4 bench shape, flat source) and 2D C5G7 with 70 groups
D=oracle/_ref/ref_driver
L3="--model simple-lattice --dims 3 --groups70 --azim 32 --spacing 0.05 --polar 6 --zspacing 0.25 --formation explicit --tol 1e-30 --max-iters 5 --quiet --no-fluxes"
for g in auto 1 2 7; do
  if [ $g = auto ]; then unset B200_GPL; else export B200_GPL=$g; fi
  $D $L3 --solver b200 --json /tmp/r.json > /tmp/r.log 2>&1
  python -c "import json; d=json.load(open('/tmp/r.json')); print('3D-70g flat GPL=$g: k %.9f' % d['keff'], 'sweep %.3f ms/iter' % (1e3*d['sweep_time_s']/d['iterations']), '%.3e integrations/s' % (d['integrations']/d['sweep_time_s']))" || tail -3 /tmp/r.log
done
for g in auto 1 2 7; do
  if [ $g = auto ]; then unset B200_GPL; else export B200_GPL=$g; fi
  echo "2D c5g7 70g GPL=$g: $(python tools/sweep_tune.py --azim 64 --spacing 0.04 --groups70 --sweeps 10 2>&1 | tail -1)"
done
