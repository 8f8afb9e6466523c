#!/bin/bash
# tally-replica sweep (B200_PHI_REPLICAS) on the 3D 70-group lattice (flat and linear source),
# the 2D simple lattice and the 2D C5G7 deck
D=oracle/_ref/ref_driver
L3="--model simple-lattice --dims 3 --groups70 --azim 32 --spacing 0.05 --polar 6 --zspacing 0.25 --formation explicit --tol 1e-30 --max-iters 5 --quiet --no-fluxes"
for sv in b200 b200ls; do for R in 1 4 16 64 auto; do
  if [ $R = auto ]; then unset B200_PHI_REPLICAS; else export B200_PHI_REPLICAS=$R; fi
  $D $L3 --solver $sv --json /tmp/r.json > /tmp/r.log 2>&1
  python -c "import json; d=json.load(open('/tmp/r.json')); print('3D-70g $sv R=$R: k %.9f' % d['keff'], 'sweep %.3f ms/iter' % (1e3*d['sweep_time_s']/d['iterations']), '%.3e integrations/s' % (d['integrations']/d['sweep_time_s']))" || tail -3 /tmp/r.log
done; done
for R in 1 4 16 64 auto; do
  if [ $R = auto ]; then unset B200_PHI_REPLICAS; else export B200_PHI_REPLICAS=$R; fi
  echo "simple-lattice 2D R=$R: $(python tools/sweep_tune.py --model simple-lattice --azim 128 --spacing 0.01 --sweeps 30 2>&1 | tail -1)"
done
for R in 1 2 4 8; do
  export B200_PHI_REPLICAS=$R
  echo "c5g7-2d R=$R: $(python tools/sweep_tune.py --azim 128 --spacing 0.02 --sweeps 20 2>&1 | tail -1)"
done
