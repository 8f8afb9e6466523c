#!/bin/bash
# timing of the device CMFD against the host Cmfd; fused CMFD loop against CPUSolver in separate processes
D=oracle/_ref/ref_driver
run() { echo "== $*"; timeout 900 $D "$@" --solver both --quiet 2>&1 | tail -2; }
A="--model c5g7-2d --azim 32 --spacing 0.1 --cmfd 51x51 --threads 16"
run $A
B200_HOST_CMFD=1 run $A
B200_CMFD_MODE=1 run $A
B200_CMFD_MODE=2 B200_CMFD_CLUSTER=8 run $A
B="--model c5g7-2d --dims 3 --azim 4 --polar 2 --spacing 1.0 --zspacing 5 --axial 9 --formation otf-stacks --cmfd 51x51x9 --max-iters 20 --threads 16"
run $B
B200_HOST_CMFD=1 run $B
B200_CMFD_MODE=1 run $B
B200_CMFD_MODE=2 B200_CMFD_CLUSTER=8 run $B
echo "== fused"
for m in "--model simple-lattice --azim 8 --spacing 0.05 --cmfd 4x4" "--model c5g7-2d --azim 8 --spacing 0.2 --cmfd 51x51 --max-iters 60"; do
  $D $m --solver cpu --threads 4 --quiet --json /tmp/cpu.json > /dev/null 2>&1
  $D $m --solver b200-fused --quiet --json /tmp/gpu.json 2>&1 | tail -2
  python - <<'PY'
import json
a=json.load(open('/tmp/cpu.json')); b=json.load(open('/tmp/gpu.json'))
import numpy as np
fa=np.array(a['fluxes']); fb=np.array(b['fluxes'])
print("fused: iters cpu %d gpu %d, dk %.3e pcm, flux err %.3e, total cpu %.3f s gpu %.3f s" % (a['iterations'], b['iterations'], abs(a['keff']-b['keff'])*1e5, np.max(np.abs(fa-fb)/np.abs(fa)), a['total_time_s'], b['total_time_s']))
PY
done
