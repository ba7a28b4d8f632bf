"""Driver loop of the GPU coupling path with the reference orchestrator's call order and config
names (splib/splib.py: read_config :436, initialize :97, run :215, step :267, finalize :410).

Only the coupling path is rebuilt: the models are the GPU stand-ins of spdummy.py (`gcm_type` /
`les_type` "gpu"); OMUSE model creation, restarts, spin-up scheduling, netCDF and MPI colour
splitting stay in the reference. One coupled step (splib.py:267-352):

    gcm.evolve_model_until_cloud_scheme(); gcm.evolve_model_cloud_scheme()
    spcpl.gather_gcm_data(...)                      -> one packed H2D copy
    spcpl.set_les_forcings  for every LES           -> K2, one launch      (reference: serial loop A)
    step_les_models(...)  + get_les_profiles        -> LES stand-in, K1
    spcpl.set_gcm_tendencies for every LES          -> K3 (+ all_gather), one D2H copy (loop B)
    gcm.evolve_model_from_cloud_scheme()
    timing.txt line: start gcm1 gather set_les_forcings set_gcm_tendencies gcm2 les...
"""
import logging
import time

import numpy as np
import torch

from . import spcpl, spio
from .spdummy import gpu_gcm, gpu_les_batch

log = logging.getLogger(__name__)

# module-global configuration with the reference's names and defaults (splib.py:39-70)
gcm_type = "gpu"
les_type = "gpu"
gcm_steps = 10
gcm_num_procs = 1
les_num_procs = 1
max_num_les = -1
les_forcing_factor = 1.0
gcm_forcing_factor = 1.0
les_spinup = 0
les_spinup_steps = 1
les_spinup_forcing_factor = 1.0
cplsurf = False
qt_forcing = "sp"
conservative_coarsening = False
variability_nudge_constant_T = False
output_name = "spifs.npz"
# GPU-path extras
les_nx, les_ny, les_nk, les_dz = 64, 64, 160, 25.0
gcm_nlev = 91
dtype = "f32"
per_column = False          # True: drive the kernels through the per-LES reference-shaped calls
gather_mode = "nccl"        # multi-GPU tendency gather: "nccl" | "p2p" | "p2p-owner" | "host" (pipeline.py)
write_diagnostics = False

gcm_model = None
les_batch = None
les_models = []
profiles = {}
firststep = True
timing_file = None
timing_rows = []


def read_config(config):
    """splib.py:436-456: dict (or JSON file name) -> module globals."""
    if isinstance(config, str):
        import json
        with open(config) as f:
            config = json.load(f)
    for key, value in (config or {}).items():
        if value is not None:
            globals()[key] = value


def initialize(config=None, geometries=None, output_geometries=None, device=None):
    """splib.py:97-212 for the GPU stand-ins: create the GCM, pick the SP columns, create one LES
    per column (all in one HBM-resident batch), attach grid_index / zf_cache / zh_cache, build the
    initial LES state from the converted GCM profiles."""
    global gcm_model, les_batch, les_models, firststep, profiles, timing_rows
    read_config(config)
    tdt = torch.float32 if dtype == "f32" else torch.float64
    ndt = np.float32 if dtype == "f32" else np.float64
    gcm_model = gpu_gcm(gcm_num_procs, nlev=gcm_nlev, dtype=ndt)
    npts = gcm_model.num_lons * gcm_model.num_lats
    if geometries is not None:
        grid_indices = [int(i) for i in geometries]
    else:
        n = npts if max_num_les is None or max_num_les < 0 else min(max_num_les, npts)
        grid_indices = list(range(n))
    for i in grid_indices:
        gcm_model.set_mask(i)                                                     # splib.py:121-122
    # multi-GPU (one process per GPU under torch.distributed): every rank owns a contiguous block of
    # the SP columns; rank 0 owns the GCM (its profiles are scattered, the tendencies gathered back)
    world, rank = 1, 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
        if len(grid_indices) % world:
            raise ValueError("%d SP columns do not divide over %d ranks" % (len(grid_indices), world))
    all_grid_indices = grid_indices
    from .pipeline import shard_columns
    lo, hi = shard_columns(len(all_grid_indices), world, rank)
    grid_indices = all_grid_indices[lo:hi]
    les_batch = gpu_les_batch(len(grid_indices), gcm_nlev, les_nx, les_ny, les_nk, les_dz, tdt, device,
                              col0=lo, couple_surface=cplsurf, gather=gather_mode,
                              ncol_total=len(all_grid_indices))
    les_batch.all_grid_indices = all_grid_indices
    les_batch.is_gcm_owner = (rank == 0)
    les_models = les_batch.models
    for les, gi in zip(les_models, grid_indices):                                 # splib.py:146-154
        les.grid_index = gi
        les.lat, les.lon = gcm_model.latitudes[gi], gcm_model.longitudes[gi]
        les.zh_cache, les.zf_cache = les.get_zh(), les.get_zf()
    spio.init_netcdf(output_name, gcm_model, les_models, gcm_model.get_start_datetime())
    gcm_model.evolve_model_until_cloud_scheme()                                   # splib.py:186-188
    gcm_model.evolve_model_cloud_scheme()
    gcm_model.first_half_step_done = True
    _gather(True)                                                                 # splib.py:196
    d = les_batch.cpl.gcm_to_les(les_batch.pipe.gcm, les_batch.pipe.zf, les_batch.pipe.zh, None, None, 1.0, 1.0,
                                 False, want_state=True)                          # convert_profiles, splib.py:202-203
    les_batch.initialize_state(d)                                                 # set_les_state,    splib.py:204
    les_batch.aux["PS"].copy_(d["ps"])
    firststep, profiles, timing_rows = True, {}, []
    if les_spinup > 0:                                                            # splib.py:206-207
        run_spinup(les_spinup, les_spinup_steps)
    return les_models


def _gather(couple_surface):
    """spcpl.gather_gcm_data on one GPU; scatter from the GCM-owning rank when the columns are sharded."""
    if les_batch.pipe.world > 1:
        spcpl.gather_gcm_data_sharded(gcm_model, les_batch, couple_surface)
    else:
        spcpl.gather_gcm_data(gcm_model, les_models, couple_surface, None, write=write_diagnostics)


def step_les_models(model_time, offset=None):
    """splib.py:554-617: advance every LES to model_time (+ the spin-up offset of the LES clocks), then fetch
    its profiles."""
    start = time.time()
    les_batch.evolve(model_time + (les_spinup if offset is None else offset))
    if per_column:
        prof = {}
        for les in les_models:
            p = spcpl.get_les_profiles(les, True)
            prof[les] = {k: v.result() for k, v in p.items()}
    else:
        prof = spcpl.get_les_profiles_all(les_batch)
    return [time.time() - start] * len(les_models), prof


def step():
    """One GCM time step (splib.py:267-352)."""
    global firststep, profiles
    t = gcm_model.get_model_time()
    delta_t = gcm_model.get_timestep()
    starttime = time.time()
    gcm1 = -time.time()
    if gcm_model.first_half_step_done:
        gcm_model.first_half_step_done = False
    else:
        gcm_model.evolve_model_until_cloud_scheme()
        gcm_model.evolve_model_cloud_scheme()
    gcm1 += time.time()
    gcm_model.step += 1
    if write_diagnostics:
        spio.update_time(t + delta_t)

    gather = -time.time()
    _gather(cplsurf)                                                              # splib.py:312
    gather += time.time()

    forc = -time.time()
    if per_column:                                                                # splib.py:317-324
        for les in les_models:
            spcpl.set_les_forcings(les, gcm_model, True, firststep, profiles.get(les, {}), dt_gcm=delta_t,
                                   factor=les_forcing_factor, couple_surface=cplsurf, qt_forcing=qt_forcing,
                                   write=write_diagnostics,
                                   variability_nudge_constant_T=variability_nudge_constant_T)
    else:
        spcpl.set_les_forcings_all(les_batch, delta_t, les_forcing_factor, cplsurf, firststep, qt_forcing=qt_forcing,
                                   variability_nudge_constant_T=variability_nudge_constant_T)
    torch.cuda.synchronize()
    forc += time.time()

    les_wall_times, profiles = step_les_models(t + delta_t)                       # splib.py:327

    tend = -time.time()
    if per_column:                                                                # splib.py:330-332
        for les in les_models:
            spcpl.set_gcm_tendencies(gcm_model, les, profiles[les], dt_gcm=delta_t, factor=gcm_forcing_factor,
                                     write=write_diagnostics, conservative=conservative_coarsening)
    else:
        if "last_forcings" not in les_batch.__dict__:
            spcpl.set_les_forcings_all(les_batch, delta_t, les_forcing_factor, cplsurf, False)
        spcpl.set_gcm_tendencies_all(gcm_model, les_batch, delta_t, gcm_forcing_factor, conservative_coarsening)
    torch.cuda.synchronize()
    tend += time.time()

    gcm2 = -time.time()
    gcm_model.evolve_model_from_cloud_scheme()                                    # splib.py:335
    gcm2 += time.time()
    row = (starttime, gcm1, gather, forc, tend, gcm2)
    timing_rows.append(row)
    if timing_file:                                                               # splib.py:340-343
        timing_file.write(('%10.2f %6.2f %6.2f %6.2f %6.2f %6.2f' % row) + ' ' +
                          ' '.join(['%6.2f' % w for w in les_wall_times[:8]]) + '\n')
        timing_file.flush()
    firststep = False
    return row


def step_spinup(spinup_length):
    """One spin-up iteration (splib.py:355-402): the LES are nudged towards the initial GCM profiles over
    `spinup_length` seconds with `les_spinup_forcing_factor`; the GCM does not step and gets no tendencies.
    Same kernels as a coupled step: K2 with dt = spinup_length, LES stand-in, K1 (the profiles feed the next
    iteration's forcings and the diagnostics store)."""
    global firststep, profiles
    if not les_models:
        return None
    starttime = time.time()
    t_les = les_batch.model_time
    forc = -time.time()
    if per_column:                                                                # splib.py:375-382
        for les in les_models:
            spcpl.set_les_forcings(les, gcm_model, True, firststep, profiles.get(les, {}), dt_gcm=spinup_length,
                                   factor=les_spinup_forcing_factor, couple_surface=cplsurf, qt_forcing=qt_forcing,
                                   write=write_diagnostics)
    else:
        spcpl.set_les_forcings_all(les_batch, spinup_length, les_spinup_forcing_factor, cplsurf, firststep,
                                   qt_forcing=qt_forcing)                        # splib.py:375-382 (constant_T stays False)
    torch.cuda.synchronize()
    forc += time.time()
    les_wall_times, profiles = step_les_models(t_les + spinup_length, offset=0)   # splib.py:386
    tend = -time.time()                                                           # profile writing, splib.py:387-391
    if write_diagnostics:
        for les in les_models:
            spcpl.write_les_profiles(les)
    tend += time.time()
    firststep = False
    row = (starttime, 0.0, 0.0, forc, tend, 0.0)                                  # splib.py:396-400
    timing_rows.append(row)
    if timing_file:
        timing_file.write(('%10.2f %6.2f %6.2f %6.2f %6.2f %6.2f' % row) + ' ' +
                          ' '.join(['%6.2f' % w for w in les_wall_times[:8]]) + '\n')
        timing_file.flush()
    return row


def run_spinup(spinup_length, spinup_steps=1):
    """splib.py:233-251: `spinup_steps` iterations covering `spinup_length` seconds in total."""
    iteration_length = spinup_length / spinup_steps
    for s in range(spinup_steps):
        if s == spinup_steps - 1:
            iteration_length = spinup_length - (spinup_steps - 1) * iteration_length
        step_spinup(iteration_length)
    log.info('  ---- Spinup done ---')


def open_timing_file(name="timing.txt"):
    """splib.py:254-262."""
    global timing_file
    timing_file = open(name, "a")
    timing_file.write('# start gcm1 gather set_les_forcings set_gcm_tendencies gcm2 ' +
                      ' '.join(str(les.grid_index) for les in les_models[:8]) + '\n# timing data\n')


def run(nsteps=None):
    """splib.py:215-230."""
    n = gcm_steps if nsteps is None else nsteps
    for _ in range(n):
        step()


def finalize():
    """splib.py:410-433."""
    global timing_file
    if write_diagnostics:
        spio.close()
    if timing_file:
        timing_file.close()
        timing_file = None
