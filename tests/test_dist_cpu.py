"""world_size-2 gloo test of the multi-GPU host logic (column sharding + tendency all_gather),
run on CPU: each rank computes its shard with the oracle, the gathered block must equal the
unsharded answer, and synthetic data must not depend on the sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import numpy_batched as nb
import synth_les
from sp_coupler_b200 import synth
from sp_coupler_b200.constants import TENDENCIES
from sp_coupler_b200.pipeline import gather_tendencies, shard_columns

NCOL, NLEV, NK, NX = 6, 19, 20, 8


def _tendencies(col0, ncol):
    zf, zh = synth.les_grid(NK, 200.0)
    gcm = synth.make_gcm_columns(ncol, NLEV, seed=5, col0=col0, ncol_total=NCOL)
    aux = synth.make_les_aux(ncol, NK, seed=5, col0=col0, ncol_total=NCOL)
    vols = synth_les.make_les_volumes(gcm, zf, NX, NX, seed=5, dtype=np.float32, col0=col0)
    r = nb.coupling_step(gcm, zf, zh, vols, aux, aux["PS"], 900.0, 1.0, 1.0, True)
    return np.stack([r["tendencies"][k] for k in TENDENCIES], axis=1), vols


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_columns(NCOL, world, rank)
    local, _ = _tendencies(lo, hi - lo)
    tend_all = torch.zeros((NCOL, 7, NLEV), dtype=torch.float64)
    gather_tendencies(torch.from_numpy(local).contiguous(), tend_all)
    if rank == 0:
        np.save(out, tend_all.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_columns_partition():
    for n, w in ((2048, 8), (10, 3), (5, 8), (0, 2)):
        parts = [shard_columns(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


def test_sharded_gather_equals_unsharded(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    full, vols = _tendencies(0, NCOL)
    assert np.array_equal(np.load(out), full)
    # data of a column depends only on its global index
    _, v1 = _tendencies(3, 3)
    assert all(np.array_equal(v1[f], vols[f][3:]) for f in vols)


def _scatter_worker(rank, world, port, outdir):
    from sp_coupler_b200.pipeline import GcmScatter, GcmStaging
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ncol = NCOL // world
    st = GcmStaging(ncol, NLEV, torch.float64, "cpu", pin=False)
    sc = GcmScatter(st, world, rank, owner=0, device="cpu", pin=False)
    if rank == 0:
        sc.fill_host(synth.make_gcm_columns(NCOL, NLEV, seed=5))
    dev = sc.scatter()
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), **{k: v.numpy() for k, v in dev.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_gcm_scatter_delivers_each_ranks_columns(tmp_path):
    """Multi-GPU gather_gcm_data: the GCM-owning rank packs per-rank blocks, one scatter delivers them."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_scatter_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    full = synth.make_gcm_columns(NCOL, NLEV, seed=5)
    for rank in range(2):
        got = np.load(str(tmp_path / ("rank%d.npz" % rank)))
        lo, hi = shard_columns(NCOL, 2, rank)
        for k in full:
            assert np.array_equal(got[k], full[k][lo:hi]), (rank, k)


class _OraclePipe(object):
    """Stand-in for CouplingPipeline on CPU tensors (what HostExchange touches): step() = the oracle on this rank's
    staged columns, for whatever level window the exchange announced."""

    cpl = None
    _zf_top = float(synth.les_grid(NK, 200.0)[0][-1])

    def __init__(self, staging, col0):
        self.staging, self.ncol, self.col0 = staging, staging.ncol, col0
        self.epoch = 0
        self._upload_src = None
        self.set_levels(NLEV)

    def upload_source(self, src):
        self._upload_src = src

    def _upload(self):
        n = self.staging.numel
        self.staging.dev_buf[:n].copy_(self._upload_src[:n])

    def set_levels(self, nlw):
        self.staging.set_levels(nlw)
        self.nlw = nlw
        self.tend = torch.zeros((self.ncol, 7, nlw), dtype=torch.float64)

    def step(self, dt, f_les, f_gcm, upload=False):
        if upload:
            self._upload()
        zf, zh = synth.les_grid(NK, 200.0)
        gcm = {k: v.numpy() for k, v in self.staging.dev.items()}
        aux = synth.make_les_aux(self.ncol, NK, seed=5, col0=self.col0, ncol_total=NCOL)
        ref = synth.make_gcm_columns(self.ncol, NLEV, seed=5, col0=self.col0, ncol_total=NCOL)
        vols = synth_les.make_les_volumes(ref, zf, NX, NX, seed=5, dtype=np.float32, col0=self.col0)
        r = nb.coupling_step(gcm, zf, zh, vols, aux, aux["PS"], dt, f_les, f_gcm, True)
        self.tend.copy_(torch.from_numpy(np.stack([r["tendencies"][k] for k in TENDENCIES], axis=1)))
        return r["forcings"]


def _exchange_worker(rank, world, port, outdir, window):
    from sp_coupler_b200.pipeline import GcmStaging, HostExchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ncol = NCOL // world
    st = GcmStaging(ncol, NLEV, torch.float64, "cpu", pin=False)
    pipe = _OraclePipe(st, rank * ncol)
    ex = HostExchange(pipe, world, rank, owner=0, register=False, tag="cputest", timeout_s=120.0, window=window)
    full = synth.make_gcm_columns(NCOL, NLEV, seed=5)
    for it in range(2):
        if rank == 0:
            g = dict(full)
            if it == 1:
                g["T"] = full["T"] + 1.5          # the host GCM moved on: every rank must see the new profiles
            ex.fill_inputs(g)
        _, out, lev0 = ex.step(900.0, 1.0, 1.0)
        if rank == 0:
            np.save(os.path.join(outdir, "out%d.npy" % it), out.numpy().copy())
            np.save(os.path.join(outdir, "lev%d.npy" % it), np.array(lev0))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("window", [False, True])
def test_host_exchange_shared_buffer(tmp_path, window):
    """Sharded host-to-host step: the GCM owner publishes every rank's inputs in one shared host buffer, each rank
    stages its own block, computes, and its tendency block lands in the shared buffer; the owner ends up with all
    columns. With the level window only the GCM levels up to the first one above the LES top travel either way; the
    block equals the full-level answer from lev0 on and everything above is zero."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_exchange_worker, args=(2, port, str(tmp_path), window), nprocs=2, join=True)
    zf, zh = synth.les_grid(NK, 200.0)
    full = synth.make_gcm_columns(NCOL, NLEV, seed=5)
    aux = synth.make_les_aux(NCOL, NK, seed=5)
    vols = synth_les.make_les_volumes(full, zf, NX, NX, seed=5, dtype=np.float32)
    for it in range(2):
        g = dict(full)
        if it == 1:
            g["T"] = full["T"] + 1.5
        r = nb.coupling_step(g, zf, zh, vols, aux, aux["PS"], 900.0, 1.0, 1.0, True)
        want = np.stack([r["tendencies"][k] for k in TENDENCIES], axis=1)
        lev0 = int(np.load(str(tmp_path / ("lev%d.npy" % it))))
        assert (lev0 > 0) == window
        assert lev0 == max(int(r["tendencies"]["start_index"].min()) - 1, 0) or not window
        assert np.array_equal(np.load(str(tmp_path / ("out%d.npy" % it))), want[:, :, lev0:]), it
        assert not want[:, :, :lev0].any()


def _unequal_worker(rank, world, port, outdir):
    from sp_coupler_b200.pipeline import GcmStaging, HostExchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = GcmStaging(3 + rank, NLEV, torch.float64, "cpu", pin=False)      # 3 columns on rank 0, 4 on rank 1
    try:
        HostExchange(_OraclePipe(st, 0), world, rank, owner=0, register=False, tag="uneq")
        msg = "no error"
    except ValueError as e:
        msg = str(e)
    open(os.path.join(outdir, "msg%d.txt" % rank), "w").write(msg)
    dist.barrier()
    dist.destroy_process_group()


def test_host_exchange_rejects_unequal_shards_on_every_rank(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_unequal_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        assert "same number of columns" in open(str(tmp_path / ("msg%d.txt" % rank))).read()
