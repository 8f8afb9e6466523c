#!/usr/bin/env python
"""computeEigenvalue over NCCL (one process per GPU) on a small deck, with a traceback dump if it stalls.
Run under torchrun: python -m torch.distributed.run --nproc-per-node 2 tools/nccl_probe.py [partition]"""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(35, exit=True)
import numpy as np
import torch
import torch.distributed as dist
from openmoc_b200.solver import B200Solver
from openmoc_b200.synth import make_tracks
from openmoc_b200.capi import FISSION_SOURCE

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
partition = sys.argv[1] if len(sys.argv) > 1 else "pair"
ft = make_tracks("simple-lattice", 8, 0.05)
print(f"[{rank}] tracks {ft.n_tracks} segs {ft.n_segments} partition {partition} graph {os.environ.get('B200_DIST_GRAPH', '1')}", flush=True)
s = B200Solver(ft, device=int(os.environ["LOCAL_RANK"]), process_group=dist.group.WORLD, partition=partition)
print(f"[{rank}] solver built", flush=True)
t0 = time.time()
s.computeEigenvalue(400, FISSION_SOURCE)
print(f"[{rank}] k {s.getKeff():.10f} iterations {s.getNumIterations()} in {time.time() - t0:.2f} s", flush=True)
s.close()
dist.destroy_process_group()
