#!/usr/bin/env python
"""Per-step device time of the multi-GPU iteration with partition="track" (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from openmoc_b200.solver import B200Solver, check
from openmoc_b200.synth import make_tracks
from openmoc_b200 import partition as P

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ft = make_tracks("c5g7-2d", num_azim=128, spacing=float(os.environ.get("SPACING", "0.02")))
s = B200Solver(ft, device=lr, process_group=dist.group.WORLD, partition="track")
s.useTorchStream()
s.zeroTrackFluxes(); s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
s.iterate(5)
torch.cuda.synchronize()
names = ["begin(sources+sweep)", "allreduce", "exchange", "end"]
acc = [0.0] * 4
host = [0.0] * 4
N = 20
for i in range(N):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    t = [time.perf_counter()]
    ev[0].record()
    check(s._lib.b200_iteration_begin(s._h, 1000 + i)); ev[1].record(); t.append(time.perf_counter())
    s._allreduce_scalar_flux(); ev[2].record(); t.append(time.perf_counter())
    s._exchange_boundary_fluxes(); ev[3].record(); t.append(time.perf_counter())
    check(s._lib.b200_iteration_end(s._h, 1000 + i, 0, 0)); ev[4].record(); t.append(time.perf_counter())
    torch.cuda.synchronize()
    for k in range(4):
        acc[k] += ev[k].elapsed_time(ev[k + 1]); host[k] += (t[k + 1] - t[k]) * 1e3
if rank == 0:
    print("plan: send %d recv %d slots of %d floats (%.1f MB out)" % (s._plan.n_send, s._plan.n_recv, ft.fluxes_per_track,
          s._plan.n_send * ft.fluxes_per_track * 4 / 1e6))
    for k in range(4):
        print("%-22s device %.3f ms   host %.3f ms" % (names[k], acc[k] / N, host[k] / N))
for label, fn in (("iterate(20), no host sync", lambda: s.iterate(20)),):
    s.resetSweepStats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms, n, _ = s.getSweepStats()
    if rank == 0:
        print("%s: %.3f ms/iteration, sweep kernel %.3f ms" % (label, e0.elapsed_time(e1) / 20, ms / max(n, 1)))
plan = s._plan
s._plan = None
s.resetSweepStats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
e0.record(); s.iterate(20); e1.record(); torch.cuda.synchronize()
ms, n, _ = s.getSweepStats()
if rank == 0:
    print("same without the exchange (wrong physics): %.3f ms/iteration, sweep kernel %.3f ms" % (e0.elapsed_time(e1) / 20, ms / max(n, 1)))
    print("tracks per rank %d (ghost %d), segments %d" % (s.tracks.n_tracks, (plan.n_send + 1) // 2, s.tracks.n_segments))
dist.destroy_process_group()
