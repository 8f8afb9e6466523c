#!/bin/bash
# Times the reference's own GPUSolver (src/accel/cuda, recompiled for sm_100a by oracle/Makefile)
# beside B200Solver on the same tracks, inside the reference's own process (ref_driver).
# usage: tools/refgpu_bench.sh [azim] [spacing] [iters]
A=${1:-128}; S=${2:-0.05}; N=${3:-20}
D=oracle/_ref/ref_driver
echo "== pin-cell check (reference golden: 261 iterations, k = 1.04666)"
for s in refgpu cpu b200; do $D --model pin-cell --azim 4 --spacing 0.1 --solver $s --tol 1e-5 --quiet --json /tmp/pc_$s.json > /tmp/pc_$s.log 2>&1 || { tail -5 /tmp/pc_$s.log; exit 1; }
  python -c "import json; d=json.load(open('/tmp/pc_$s.json')); print('$s', d['iterations'], '%.10f' % d['keff'])"; done
echo "== c5g7-2d azim $A spacing $S, $N iterations"
for cfg in "0 0" "2368 128" "4736 64" "9472 128" "18944 64"; do set -- $cfg
  $D --model c5g7-2d --azim $A --spacing $S --solver refgpu --gpu-blocks $1 --gpu-threads $2 --tol 1e-30 --max-iters $N --quiet --no-fluxes --json /tmp/rg.json > /tmp/rg.log 2>&1
  python -c "import json; d=json.load(open('/tmp/rg.json')); print('refgpu B=$1 T=$2: n_seg', d['n_segments'], 'k %.8f' % d['keff'], 'sweep %.4f s/iter' % (d['sweep_time_s']/d['iterations']), '%.3e integrations/s' % (d['integrations']/d['sweep_time_s']))" || tail -5 /tmp/rg.log
done
$D --model c5g7-2d --azim $A --spacing $S --solver b200 --tol 1e-30 --max-iters $N --quiet --no-fluxes --json /tmp/rg.json > /tmp/rg.log 2>&1
python -c "import json; d=json.load(open('/tmp/rg.json')); print('b200: n_seg', d['n_segments'], 'k %.8f' % d['keff'], 'sweep %.4f s/iter' % (d['sweep_time_s']/d['iterations']), '%.3e integrations/s' % (d['integrations']/d['sweep_time_s']))" || tail -5 /tmp/rg.log
