#!/usr/bin/env python
"""Sweep throughput of B200Solver on a B2TRK track file (any deck the reference can ray-trace:
3D, OTF, ...), on 1..N GPUs.  Under torchrun every rank loads the file and keeps its shard.

  oracle/_ref/ref_driver --model c5g7-2d --dims 3 --azim 16 --spacing 0.2 --polar 4 --zspacing 2.0 \\
      --formation otf-stacks --mode none --solver cpu --threads 16 --quiet --dump-tracks /tmp/c5g7_3d.b2trk
  python tools/trackfile_bench.py /tmp/c5g7_3d.b2trk [--partition chain|track|pair] [--steps 20]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/trackfile_bench.py /tmp/c5g7_3d.b2trk
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmoc_b200.solver import B200Solver
from openmoc_b200.trackfile import read_trackfile

ap = argparse.ArgumentParser()
ap.add_argument("trackfile")
ap.add_argument("--partition", default="chain", choices=["pair", "chain", "track"])
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.perf_counter()
ft = read_trackfile(args.trackfile)
t_load = time.perf_counter() - t0
t0 = time.perf_counter()
s = B200Solver(ft, device=local, process_group=(dist.group.WORLD if dist else None), partition=args.partition)
s.useTorchStream()
t_setup = time.perf_counter() - t0
s.zeroTrackFluxes(); s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
s.iterate(args.warmup)
torch.cuda.synchronize()
if dist: dist.barrier()
s.resetSweepStats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s.iterate(args.steps); e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
if dist: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
sweep_ms, n_sweeps, _ = s.getSweepStats()
W = 2.0 * ft.fluxes_per_track * ft.n_segments
if rank == 0:
    print(json.dumps({"trackfile": os.path.basename(args.trackfile), "n_gpus": world, "partition": args.partition if world > 1 else None,
                      "n_tracks": ft.n_tracks, "n_segments": ft.n_segments, "n_fsrs": ft.n_fsrs, "groups": ft.num_groups,
                      "solve_3d": ft.solve_3d, "integrations_per_s": W * args.steps / (ms.item() * 1e-3),
                      "ms_per_iteration": ms.item() / args.steps, "sweep_kernel_ms_rank0": sweep_ms / max(n_sweeps, 1),
                      "k_eff": s.getKeff(), "load_s": round(t_load, 2), "partition_and_upload_s": round(t_setup, 2)}))
if dist: dist.destroy_process_group()
