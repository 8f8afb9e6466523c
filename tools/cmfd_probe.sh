#!/bin/bash
# device CMFD next to the reference's host Cmfd (ref_driver --solver both), small decks first
D=oracle/_ref/ref_driver
run() { echo "== $*"; timeout 600 $D "$@" --solver both --quiet 2>&1 | tail -3; }
run --model simple-lattice --azim 4 --spacing 0.12 --cmfd 2x2
B200_HOST_CMFD=1 run --model simple-lattice --azim 4 --spacing 0.12 --cmfd 2x2
run --model simple-lattice --azim 8 --spacing 0.05 --cmfd 4x4
run --model simple-lattice --azim 8 --spacing 0.05 --cmfd 4x4 --no-knearest
run --model c5g7-2d --azim 8 --spacing 0.2 --cmfd 51x51 --threads 1 --max-iters 60
B200_HOST_CMFD=1 run --model c5g7-2d --azim 8 --spacing 0.2 --cmfd 51x51 --threads 1 --max-iters 60
B200_CMFD_MODE=1 run --model c5g7-2d --azim 8 --spacing 0.2 --cmfd 51x51 --threads 1 --max-iters 60
run --model simple-lattice --dims 3 --azim 4 --polar 2 --spacing 0.24 --zspacing 0.9 --cmfd 2x2x2
run --model simple-lattice --dims 3 --azim 4 --polar 2 --spacing 0.24 --zspacing 0.9 --cmfd 2x2x2 --ls --formation otf-stacks
run --model c5g7-2d --dims 3 --azim 4 --polar 2 --spacing 1.0 --zspacing 10 --formation otf-stacks --cmfd 51x51x3 --max-iters 20 --threads 4
echo "== sanitizer"
timeout 900 compute-sanitizer --tool memcheck $D --model simple-lattice --azim 4 --spacing 0.12 --cmfd 2x2 --solver b200 --quiet --max-iters 5 2>&1 | tail -5
