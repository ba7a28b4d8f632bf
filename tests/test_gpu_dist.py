"""Multi-GPU parity where the driver sees it (-m gpu): the sharded step on 2 / 4 / 8 GPUs of the box (one process per
GPU, NCCL process group) must give, bit for bit, the tendencies one GPU computes for all columns - through every
delivery route of the packed block (reference analogue: the 7 gcm.set_profile_tendency calls per column,
/root/reference/splib/spcpl.py:535-542):

  p2p-owner  K3 stores the block into the GCM owner's gather buffer over NVLink, in-kernel barrier   (eager + CUDA graph)
  p2p        K3 stores it into every rank's buffer                                                    (eager + CUDA graph)
  nccl       all_gather_into_tensor after K3
  host       HostExchange: K3 stores it into the host GCM's pinned buffer shared by all ranks, level window, flag polling

Skipped when the box has fewer GPUs than ranks (the 1-GPU round-end run); run it with `gpurun --gpus 2|8`.
"""
import os
import socket

import numpy as np
import pytest

NCOL, NX, NK, NLEV, STEPS = 6, 16, 160, 91, 4


def _worker(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    msgs = []
    try:
        import synth_les
        from sp_coupler_b200 import synth
        from sp_coupler_b200.coupler import Coupler
        from sp_coupler_b200.pipeline import CouplingPipeline, HostExchange
        cpl = Coupler(dev)
        ntot = NCOL * world
        zf, zh = synth.les_grid(NK)
        up = lambda d: {k: torch.from_numpy(v).to(dev) for k, v in d.items()}

        def inputs(col0, ncol):
            gcm = synth.make_gcm_columns(ncol, NLEV, seed=11, dtype=np.float32, col0=col0, ncol_total=ntot)
            aux = up(synth.make_les_aux(ncol, NK, seed=11, dtype=np.float32, col0=col0, ncol_total=ntot))
            vols = synth_les.device_les_volumes(cpl, gcm, zf, NX, NX, seed=11, col0=col0)
            return gcm, aux, vols

        def run(pipe, vols, aux, gcm, graph, sync_each=True, lev0=0):
            """STEPS steps; the LES state changes between steps so that a stale gather buffer would show. Without
            sync_each the steps are queued back to back (ranks may run a step ahead of each other: the alternating
            gather buffers and the in-kernel barrier must keep them apart) and only the last block is returned."""
            vols = [v.clone() for v in vols]
            pipe.stage_host(gcm, lev0=lev0)
            pipe.staging.upload()
            pipe.attach_les(vols, aux)
            pipe.les_profiles()
            if graph:
                pipe.capture(900.0, 1.0, 1.0)
            outs = []
            for it in range(STEPS):
                vols[1].mul_(1.0 + 1e-3 * (it + 1))
                vols[2].mul_(1.0 + 1e-2 * (it + 1))
                pipe.step(900.0, 1.0, 1.0)
                if sync_each or it == STEPS - 1:
                    torch.cuda.synchronize()
                    dist.barrier()
                    outs.append(pipe.tend_all.clone())
            return outs

        # one GPU, all columns: what every sharded route must reproduce
        gcm_all, aux_all, vols_all = inputs(0, ntot)
        ref = run(CouplingPipeline(cpl, zf, zh, ntot, NLEV, torch.float32, gather=False), vols_all, aux_all, gcm_all, False)
        assert not torch.equal(ref[0], ref[-1])
        gcm, aux, vols = inputs(rank * NCOL, NCOL)
        # the job-wide live level window (every rank must run the same one: the gathered block is [..][7][nlw])
        from sp_coupler_b200.pipeline import first_live_level
        lev_all = first_live_level(gcm_all["Zgfull"], gcm_all["Zghalf"][:, -1], float(zf[-1]))
        if lev_all <= 0:
            msgs.append("no level window in this case (lev0=%d)" % lev_all)
        for mode, graph, sync_each, lev0 in (("p2p-owner", False, True, 0), ("p2p-owner", True, True, 0), ("p2p-owner", False, False, 0),
                                             ("p2p-owner", True, False, 0), ("p2p", False, True, 0), ("p2p", True, False, 0),
                                             ("nccl", False, True, 0), ("p2p-owner", True, True, lev_all), ("p2p", False, False, lev_all),
                                             ("nccl", False, True, lev_all)):
            pipe = CouplingPipeline(cpl, zf, zh, NCOL, NLEV, torch.float32, gather=mode)
            outs = run(pipe, vols, aux, gcm, graph, sync_each, lev0)
            want = [r[:, :, lev0:] for r in (ref if sync_each else ref[-1:])]
            if rank == 0 or mode != "p2p-owner":
                for it, (a, b) in enumerate(zip(outs, want)):
                    if not torch.equal(a, b):
                        msgs.append("%s graph=%s sync_each=%s lev0=%d step %d: gathered block differs from the single-GPU result"
                                    % (mode, graph, sync_each, lev0, it))
            if pipe.sync_error():
                msgs.append("%s graph=%s: sync error %d" % (mode, graph, pipe.sync_error()))
            del pipe
        # host-resident GCM: shared pinned buffer, K3 stores into it
        for graph in (False, True):
            hp = CouplingPipeline(cpl, zf, zh, NCOL, NLEV, torch.float32, gather=False)
            ex = HostExchange(hp, world, rank, owner=0, tag="disttest")

            def host_step():
                if rank == 0:
                    ex.fill_inputs(gcm_all)
                _, out, lev0 = ex.step(900.0, 1.0, 1.0)
                dist.barrier()
                return (out.clone(), lev0)

            hv = [v.clone() for v in vols]
            hp.attach_les(hv, aux)
            hp.les_profiles()
            if graph:       # adopt the window first, then record
                if rank == 0:
                    ex.fill_inputs(gcm_all)
                ex.step(900.0, 1.0, 1.0)
                hp.capture(900.0, 1.0, 1.0, upload=True)      # H2D of the rank's block is the graph's first node
            for it in range(STEPS):
                hv[1].mul_(1.0 + 1e-3 * (it + 1))
                hv[2].mul_(1.0 + 1e-2 * (it + 1))
                out, lev0 = host_step()
                if rank == 0:
                    if lev0 <= 0:
                        msgs.append("host graph=%s: no level window (lev0=%d)" % (graph, lev0))
                    if not torch.equal(out, ref[it][:, :, lev0:].cpu()):
                        msgs.append("host graph=%s step %d: host block differs from the single-GPU result" % (graph, it))
                    if ref[it][:, :, :lev0].any():
                        msgs.append("host graph=%s step %d: non-zero tendency above the window" % (graph, it))
            if hp.sync_error():
                msgs.append("host graph=%s: sync error %d" % (graph, hp.sync_error()))
            torch.cuda.synchronize()
            dist.barrier()
            ex.close()
    except Exception as e:      # noqa: BLE001
        import traceback
        msgs.append("rank %d raised: %s\n%s" % (rank, e, traceback.format_exc()))
    open(os.path.join(outdir, "rank%d.txt" % rank), "w").write("\n".join(msgs) if msgs else "ok")
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:           # noqa: BLE001
        pass


@pytest.mark.gpu
@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_step_equals_single_gpu(tmp_path, world):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs on the box (has %d)" % (world, torch.cuda.device_count() if torch.cuda.is_available() else 0))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert open(str(tmp_path / ("rank%d.txt" % rank))).read() == "ok", rank


def _c5_worker(rank, world, port, outdir):
    """BASELINE.json configs[4]: 16384 columns of 32x32x160, L137, float32, sharded over 8 GPUs (2048 per rank)."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    msgs = []
    try:
        from oracle import numpy_batched as nb
        import synth_les
        from sp_coupler_b200 import synth
        from sp_coupler_b200.constants import TENDENCIES
        from sp_coupler_b200.coupler import Coupler
        from sp_coupler_b200.pipeline import CouplingPipeline
        cpl = Coupler(dev)
        ncol, nx, nk, nlev = 16384 // world, 32, 160, 137
        ntot = ncol * world
        zf, zh = synth.les_grid(nk)

        def job(r, gather):
            gcm = synth.make_gcm_columns(ncol, nlev, seed=46, dtype=np.float32, col0=r * ncol, ncol_total=ntot)
            aux = synth.make_les_aux(ncol, nk, seed=46, dtype=np.float32, col0=r * ncol, ncol_total=ntot)
            pipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, torch.float32, gather=gather)
            pipe.staging.fill_host(gcm)
            pipe.staging.upload()
            pipe.attach_les(synth_les.device_les_volumes(cpl, gcm, zf, nx, nx, seed=46, col0=r * ncol),
                            {k: torch.from_numpy(v).to(dev) for k, v in aux.items()})
            pipe.les_profiles()
            return pipe, gcm, aux

        pipe, gcm, aux = job(rank, "p2p-owner")
        pipe.capture(900.0, 1.0, 1.0)
        for _ in range(2):
            pipe.step(900.0, 1.0, 1.0)
        torch.cuda.synchronize()
        dist.barrier()
        if pipe.sync_error():
            msgs.append("sync error %d" % pipe.sync_error())
        # every rank: one of ITS columns against the CPU oracle on host-generated inputs (bit-identical generator)
        c = (37 * (rank + 1)) % ncol
        g1 = {k: v[c:c + 1] for k, v in gcm.items()}
        a1 = {k: v[c:c + 1] for k, v in aux.items()}
        hv = synth_les.make_les_volumes(g1, zf, nx, nx, seed=46, dtype=np.float32, col0=rank * ncol + c)
        ref = nb.coupling_step(g1, zf, zh, hv, a1, a1["PS"], 900.0, 1.0, 1.0, True)
        got = pipe.tend[c].cpu().numpy()
        for i, k in enumerate(TENDENCIES):
            d = np.abs(got[i] - ref["tendencies"][k][0]).max() / max(np.abs(ref["tendencies"][k][0]).max(), 1e-300)
            if d > 1e-4:
                msgs.append("rank %d column %d %s: rel err %.2e vs the oracle" % (rank, c, k, d))
        if not np.array_equal(pipe.slab["cnt"][c].cpu().numpy(), ref["cnt"][0]):
            msgs.append("rank %d column %d: cloud counts differ from the oracle" % (rank, c))
        if rank == 0:       # sharding invariance: the gathered block == the owner's own computation of every rank's columns
            gathered = pipe.tend_all.clone()
            for r in range(world):
                sp, _, _ = job(r, False)
                sp.step_device(900.0, 1.0, 1.0)
                torch.cuda.synchronize()
                if not torch.equal(sp.tend, gathered[r * ncol:(r + 1) * ncol]):
                    msgs.append("block of rank %d: gathered bits differ from the single-GPU computation" % r)
                del sp
                torch.cuda.empty_cache()
        dist.barrier()
    except Exception as e:      # noqa: BLE001
        import traceback
        msgs.append("rank %d raised: %s\n%s" % (rank, e, traceback.format_exc()))
    open(os.path.join(outdir, "rank%d.txt" % rank), "w").write("\n".join(msgs) if msgs else "ok")
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:           # noqa: BLE001
        pass


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_c5_sharded_over_8_gpus(tmp_path):
    """16384 columns over 8 GPUs: a spot column per rank against the oracle, and the block gathered on the GCM owner
    (K3's NVLink stores, graph-replayed step) equal to the owner's own single-GPU computation of every rank's columns."""
    import torch
    import torch.multiprocessing as mp
    world = 8
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs 8 GPUs on the box")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_c5_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert open(str(tmp_path / ("rank%d.txt" % rank))).read() == "ok", rank


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_sharded_driver_equals_single_gpu(tmp_path):
    """spmaster.py (the reference's driver loop, splib.step) under torch.distributed.run on 2 GPUs: after 3 coupled steps
    the GCM state of the SP columns equals the single-GPU run bit for bit, with every gather mode - NCCL all_gather,
    fused NVLink stores (to every rank / to the owner), and the shared pinned host buffer with the level window."""
    import subprocess
    import sys
    import torch
    from conftest import ROOT
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on the box")
    common = ["--steps", "3", "--numles", "16", "--nx", "16", "--ny", "16", "--cplsurf"]
    one = str(tmp_path / "one.npz")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "spmaster.py")] + common + ["--save_state", one],
                       capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr[-2000:]
    ref = np.load(one)
    for mode in ("nccl", "p2p", "p2p-owner", "host"):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        out = str(tmp_path / ("two_%s.npz" % mode))
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "spmaster.py")] +
                           common + ["--gather", mode, "--save_state", out], capture_output=True, text=True, timeout=400,
                           cwd=str(tmp_path))
        assert p.returncode == 0, (mode, p.stderr[-3000:])
        got = np.load(out)
        for k in ref.files:
            assert np.array_equal(got[k], ref[k]), (mode, k)
    assert float(np.abs(ref["T"]).sum()) > 0
