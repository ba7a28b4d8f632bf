"""Run under torchrun (N>=2): the fused K3->peer-store gather must give the same gathered block as
the NCCL all_gather, on every rank, over several steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from sp_coupler_b200 import synth
from sp_coupler_b200.coupler import Coupler
from sp_coupler_b200.pipeline import CouplingPipeline

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ncol, nlev, nx, nk = 64, 91, 32, 160
cpl = Coupler(dev)
zf, zh = synth.les_grid(nk)
gcm = synth.make_gcm_columns(ncol, nlev, dtype=np.float32, col0=rank * ncol, ncol_total=ncol * world)
aux = synth.make_les_aux(ncol, nk, dtype=np.float32, col0=rank * ncol, ncol_total=ncol * world)
vols = synth.device_les_volumes(cpl, gcm, zf, nx, nx, col0=rank * ncol)
out = {}
for mode in ("nccl", "p2p", "p2p-owner"):
    pipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, torch.float32, gather=mode)
    pipe.staging.fill_host(gcm)
    pipe.staging.upload()
    pipe.attach_les(vols, {k: torch.from_numpy(v).to(dev) for k, v in aux.items()})
    pipe.les_profiles()
    res = []
    for step in range(4):
        pipe.step_host(900.0, 1.0, 1.0 + step)     # a different forcing factor every step
        res.append(pipe.tend_all.clone())
    out[mode] = res
torch.cuda.synchronize()
ok = all(torch.equal(a, b) for a, b in zip(out["nccl"], out["p2p"]))
if rank == 0:       # gather-to-owner: only the GCM-owning rank holds every block
    ok = ok and all(torch.equal(a, b) for a, b in zip(out["nccl"], out["p2p-owner"]))
nz = float(out["p2p"][-1].abs().sum()) > 0
flag = torch.tensor([int(ok and nz)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("p2p gather == nccl all_gather on all %d ranks: %s" % (world, bool(flag.item())))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
