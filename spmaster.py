#!/usr/bin/env python
"""Thin driver for the GPU coupling path with the reference CLI's flag names for this path
(reference: spmaster.py:76-238 -> splib.initialize / run / finalize, spmaster.py:263-269).

Only the stand-in models exist here (--gcmtype gpu, --lestype gpu): OpenIFS / DALES workers,
geometry files, MPI channels and restarts belong to the reference's control plane.

    python spmaster.py --steps 5 --numles 256 --cplsurf
"""
import argparse
import logging
import sys

logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s")
log = logging.getLogger("spmaster")


def main(argv=None):
    p = argparse.ArgumentParser(description="superparameterization coupling step on B200 (stand-in models)")
    p.add_argument("--steps", dest="gcm_steps", type=int, default=5, help="number of GCM time steps")
    p.add_argument("--numles", dest="max_num_les", type=int, default=256, help="number of SP columns / LES models")
    p.add_argument("--gcmtype", dest="gcm_type", default="gpu", choices=["gpu"])
    p.add_argument("--lestype", dest="les_type", default="gpu", choices=["gpu"])
    p.add_argument("--cplsurf", dest="cplsurf", action="store_true", help="couple surface fluxes")
    p.add_argument("--lesforcingfactor", dest="les_forcing_factor", type=float, default=1.0)
    p.add_argument("--gcmforcingfactor", dest="gcm_forcing_factor", type=float, default=1.0)
    p.add_argument("--spinup", dest="les_spinup", type=float, default=0, help="time (s) of initial LES spin-up towards the initial profile")
    p.add_argument("--spinup_steps", dest="les_spinup_steps", type=int, default=1, help="number of iterations of the spin-up nudging")
    p.add_argument("--spinup_forcing", dest="les_spinup_forcing_factor", type=float, default=1.0, help="forcing strength during LES spin-up")
    p.add_argument("--qt_forcing", dest="qt_forcing", default="sp", choices=["sp", "variance"])
    p.add_argument("--conservative_coarsening", dest="conservative_coarsening", action="store_true")
    p.add_argument("--nx", dest="les_nx", type=int, default=64)
    p.add_argument("--ny", dest="les_ny", type=int, default=64)
    p.add_argument("--nk", dest="les_nk", type=int, default=160)
    p.add_argument("--nlev", dest="gcm_nlev", type=int, default=91)
    p.add_argument("--dtype", dest="dtype", default="f32", choices=["f32", "f64"])
    p.add_argument("--per_column", dest="per_column", action="store_true",
                   help="drive the kernels through the per-LES reference-shaped calls")
    p.add_argument("--output", dest="output_name", default="spifs.npz")
    p.add_argument("--write", dest="write_diagnostics", action="store_true")
    p.add_argument("--gather", dest="gather_mode", default="nccl", choices=["nccl", "p2p", "p2p-owner", "host"],
                   help="multi-GPU tendency gather (under torch.distributed.run): NCCL all_gather, fused NVLink stores, "
                        "or 'host' = every rank exchanges its own columns with the GCM through one shared pinned host buffer")
    p.add_argument("--save_state", default=None, help="write the final GCM state of the SP columns to this .npz")
    args = p.parse_args(argv)

    import os
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:   # one process per GPU: columns are sharded, rank 0 owns the GCM
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))

    from sp_coupler_b200 import splib
    cfg = {k: v for k, v in args.__dict__.items() if k != "save_state"}
    splib.initialize(cfg)
    if rank == 0:
        splib.open_timing_file()
    splib.run(args.gcm_steps)
    splib.finalize()
    if args.save_state and rank == 0:
        import numpy as np
        cols = splib.les_batch.all_grid_indices
        np.savez(args.save_state, **{k: splib.gcm_model.state[k][cols] for k in ("T", "SH", "QL", "QI", "U", "V", "A")})
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    rows = splib.timing_rows
    if rows and rank == 0:
        n = len(splib.les_models)
        f = sum(r[3] for r in rows[1:]) / max(len(rows) - 1, 1)
        t = sum(r[4] for r in rows[1:]) / max(len(rows) - 1, 1)
        log.info("%d columns, %d steps: set_les_forcings %.2f ms/step, set_gcm_tendencies(+slab) %.2f ms/step",
                 n, len(rows), f * 1e3, t * 1e3)
    return 0


if __name__ == "__main__":
    sys.exit(main())
