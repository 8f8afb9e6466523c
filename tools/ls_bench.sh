#!/bin/bash
# linear-source sweep throughput: 2D C5G7 (128 azim, 0.05 cm) and the 3D 70-group lattice (C4 bench shape)
D=oracle/_ref/ref_driver
run() { $D "$@" --tol 1e-30 --quiet --no-fluxes --json /tmp/r.json > /tmp/r.log 2>&1
  python -c "import json; d=json.load(open('/tmp/r.json')); print(d['model'], d['dims'], d['solver'], d['n_segments'], 'k %.9f' % d['keff'], 'sweep %.3f ms/iter' % (1e3*d['sweep_time_s']/d['iterations']), '%.3e integrations/s' % (d['integrations']/d['sweep_time_s']))" || tail -3 /tmp/r.log; }
run --model c5g7-2d --azim 128 --spacing 0.05 --solver b200ls --max-iters 10
run --model simple-lattice --dims 3 --groups70 --azim 32 --spacing 0.05 --polar 6 --zspacing 0.25 --formation explicit --solver b200ls --max-iters 5
