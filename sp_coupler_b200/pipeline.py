"""Device-resident batched coupling step for the columns one rank owns.

This is what the `splib.step`-shaped driver (sp_coupler_b200/splib.py) and bench.py run per GCM
time step instead of the reference's two serial Python loops over LES models
(splib/splib.py:317-323 and :330-332):

    H2D  GCM profiles of the step (one packed copy from pinned host memory)   <- gather_gcm_data
    K2   gcm_to_les   forcings on the LES from the previous slab means        <- set_les_forcings
         [ the LES models time-step here; external to the coupling path ]
    K1   slab_reduce  slab means + cloud mask of the LES volumes              <- get_les_profiles
    K3   les_to_gcm   tendencies on the GCM, packed [ncol][7][nlev]           <- set_gcm_tendencies
    NCCL all_gather of the packed tendency block when columns are sharded
    D2H  tendencies to the rank that owns the GCM

Columns are independent, so ranks own contiguous column blocks and the only exchange is the
tendency gather.
"""
import numpy as np
import torch

from .constants import surf_vars
from .coupler import GCM_FULL, GCM_HALF


def shard_columns(ncol_total, world_size, rank):
    """Contiguous block partition: rank r owns [lo, hi). Remainder columns go to the first ranks."""
    base, rem = divmod(ncol_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_tendencies(tend_local, tend_all, group=None):
    """The path's only exchange: all_gather of the packed [ncol_local][7][nlev]
    tendency blocks into [ncol_local*world][7][nlev] so the rank that owns the GCM holds every
    column's tendencies (reference analogue: 7 set_profile_tendency RPCs per column,
    spcpl.py:535-542). NCCL on device tensors, gloo on CPU tensors (tests)."""
    torch.distributed.all_gather_into_tensor(tend_all, tend_local, group=group)
    return tend_all


class GcmStaging(object):
    """Packed struct-of-arrays staging of the GCM inputs (gcm_vars + surf_vars, spcpl.py:32-33):
    one pinned host buffer, one device buffer, one async copy per step."""

    def __init__(self, ncol, nlev, dtype, device, pin=True):
        self.ncol, self.nlev, self.dtype = ncol, nlev, dtype
        sizes = [(n, (ncol, nlev)) for n in GCM_FULL] + [(n, (ncol, nlev + 1)) for n in GCM_HALF] + \
                [(n, (ncol,)) for n in surf_vars]
        total = sum(int(np.prod(s)) for _, s in sizes)
        self.host_buf = torch.empty(total, dtype=dtype, pin_memory=pin and torch.cuda.is_available())
        self.dev_buf = torch.empty(total, dtype=dtype, device=device)
        self.host, self.dev = {}, {}
        off = 0
        for n, s in sizes:
            cnt = int(np.prod(s))
            self.host[n] = self.host_buf[off:off + cnt].view(*s)
            self.dev[n] = self.dev_buf[off:off + cnt].view(*s)
            off += cnt
        self.nbytes = total * self.host_buf.element_size()

    def fill_host(self, gcm):
        """Copy a dict of numpy / CPU-tensor arrays into the pinned buffer (what the host GCM does)."""
        for n, h in self.host.items():
            if n in gcm:
                h.copy_(torch.as_tensor(np.ascontiguousarray(gcm[n])) if not isinstance(gcm[n], torch.Tensor) else gcm[n])

    def upload(self):
        self.dev_buf.copy_(self.host_buf, non_blocking=True)
        return self.dev


class GcmScatter(object):
    """Multi-GPU form of gather_gcm_data (spcpl.py:55-86): the rank that owns the GCM packs every
    rank's GcmStaging block into one pinned buffer, uploads it once, and one scatter over the process
    group (NCCL / NVLink on GPUs, gloo on CPU tensors in the tests) delivers each rank's block straight
    into its staging device buffer. Equal column counts per rank."""

    def __init__(self, staging, world, rank, owner=0, group=None, device=None, pin=True):
        self.staging, self.world, self.rank, self.owner, self.group = staging, world, rank, owner, group
        self.per_rank = staging.dev_buf.numel()
        self.host_all = self.dev_all = None
        if rank == owner:
            self.host_all = torch.empty(world * self.per_rank, dtype=staging.dtype,
                                        pin_memory=pin and torch.cuda.is_available())
            self.dev_all = torch.empty(world * self.per_rank, dtype=staging.dtype,
                                       device=device if device is not None else staging.dev_buf.device)

    def fill_host(self, gcm_all):
        """Owner only. gcm_all: dict of [world*ncol, ...] host arrays in global column order."""
        st, ncol = self.staging, self.staging.ncol
        for r in range(self.world):
            base = r * self.per_rank
            off = 0
            for n, h in st.host.items():
                cnt = h.numel()
                src = gcm_all[n][r * ncol:(r + 1) * ncol]
                src = src if isinstance(src, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(src))
                self.host_all[base + off:base + off + cnt].view(h.shape).copy_(src)
                off += cnt

    def scatter(self):
        """All ranks. Returns the rank's device views (staging.dev)."""
        if self.rank == self.owner:
            self.dev_all.copy_(self.host_all, non_blocking=True)
            chunks = list(self.dev_all.chunk(self.world))
            torch.distributed.scatter(self.staging.dev_buf, scatter_list=chunks, src=self.owner, group=self.group)
        else:
            torch.distributed.scatter(self.staging.dev_buf, src=self.owner, group=self.group)
        return self.staging.dev


def bind_host_thread_to_gpu(device):
    """Restricts this process to the CPUs NVML reports as local to `device` (its NUMA node), so that the pinned host
    pages it first touches and its PCIe copies stay on the GPU's side of the socket interconnect. Returns the CPU
    list, or None when NVML / the affinity call is unavailable (then nothing changes)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(device)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08X:%02X:%02X.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:          # noqa: BLE001 - an optimisation only
        return None


class HostExchange(object):
    """Host side of a sharded step when the GCM lives in HOST memory of one process (OpenIFS does): one pinned
    host buffer shared by all ranks of the node (a /dev/shm mapping that every rank registers with CUDA).

        GCM owner   writes every rank's packed input block into `inp[r]`, then publishes the step number
        every rank  waits for it, copies ITS block host->device, runs the step on its columns, copies ITS
                    tendency block device->host into `out[r*ncol:(r+1)*ncol]`, publishes "done"
        GCM owner   waits for all ranks: `out` holds [world*ncol][7][nlev], no device gather, no collective

    Every rank owns the same number of columns (as with GcmScatter); the constructor checks that collectively,
    so the copies use every GPU's PCIe link at once instead of funnelling world*ncol columns through the
    owner GPU's link (reference analogue: the master gathers every profile over its own channels,
    spcpl.py:55-86, 535-542). Flags are int64 words in the same mapping (single writer each, monotonic).
    Works on CPU tensors too (no registration) for the gloo tests."""

    FLAG_WORDS = 64

    def __init__(self, staging, world, rank, owner=0, group=None, register=True, tag="x", timeout_s=60.0):
        import os
        self.staging, self.world, self.rank, self.owner, self.group = staging, world, rank, owner, group
        self.timeout_s = timeout_s
        ncol, nlev, dtype = staging.ncol, staging.nlev, staging.dtype
        esize = torch.empty((), dtype=dtype).element_size()
        self.per_rank_in = staging.host_buf.numel()
        self.per_rank_out = ncol * 7 * nlev
        nin = world * self.per_rank_in * esize
        nout = world * self.per_rank_out * esize
        al = lambda n: (n + 4095) // 4096 * 4096
        self._off_out = al(self.FLAG_WORDS * 8)
        self._off_in = self._off_out + al(nout)
        self.nbytes = self._off_in + al(nin)
        path = "/dev/shm/spcpl_b200_%s_%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "r"), tag)
        # every phase ends in an all_reduce(MIN) of "it worked here", so that all ranks either get the buffer or raise
        # together - a rank that fails alone must never leave the others waiting in a collective
        dev = "cuda" if torch.distributed.get_backend(group) == "nccl" else "cpu"
        shape = torch.tensor([ncol, -ncol, nlev, -nlev], device=dev)
        torch.distributed.all_reduce(shape, op=torch.distributed.ReduceOp.MIN, group=group)
        if int(shape[0]) != -int(shape[1]) or int(shape[2]) != -int(shape[3]):
            raise ValueError("HostExchange needs the same number of columns and levels on every rank "
                             "(columns %d..%d, levels %d..%d)" % (int(shape[0]), -int(shape[1]), int(shape[2]), -int(shape[3])))
        err = None
        try:
            if rank == owner:
                with open(path, "wb") as f:
                    f.truncate(self.nbytes)
        except Exception as e:          # noqa: BLE001
            err = e
        ok = self._agree(err is None, group)
        self.raw, self.registered = None, False
        if ok:
            try:
                self.raw = torch.from_file(path, shared=True, size=self.nbytes, dtype=torch.uint8)
                # first touch: this rank's input and output blocks are allocated on the NUMA node it runs on
                self.raw[self._off_in + rank * self.per_rank_in * esize:self._off_in + (rank + 1) * self.per_rank_in * esize].zero_()
                self.raw[self._off_out + rank * self.per_rank_out * esize:self._off_out + (rank + 1) * self.per_rank_out * esize].zero_()
                if register and torch.cuda.is_available():
                    rc = torch.cuda.cudart().cudaHostRegister(self.raw.data_ptr(), self.nbytes, 0)
                    if int(rc) != 0:
                        raise RuntimeError("cudaHostRegister failed (%s)" % (rc,))
                    self.registered = True
            except Exception as e:      # noqa: BLE001
                err = e
            ok = self._agree(err is None, group)
        if rank == owner:
            try:
                os.unlink(path)        # the mapping stays alive in every process; nothing is left behind
            except OSError:
                pass
        if not ok:
            self.close()
            raise RuntimeError("HostExchange: shared pinned host buffer unavailable on at least one rank (%s)" % (err,))
        self.flags = self.raw[:self.FLAG_WORDS * 8].view(torch.int64)          # [0] inputs ready, [1+r] rank r done
        self.out = self.raw[self._off_out:self._off_out + nout].view(dtype).view(world * ncol, 7, nlev)
        self.inp = self.raw[self._off_in:self._off_in + nin].view(dtype).view(world, self.per_rank_in)
        if rank == owner:
            self.flags.zero_()
        torch.distributed.barrier(group=group)
        self.step_no = 0

    @staticmethod
    def _agree(ok, group):
        dev = "cuda" if torch.distributed.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor([1 if ok else 0], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=group)
        return bool(int(t.item()))

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.raw.data_ptr())
            self.registered = False

    def fill_inputs(self, gcm_all):
        """Owner only. gcm_all: dict of [world*ncol, ...] host arrays in global column order (what
        gather_gcm_data fetched from the host GCM), packed per rank in GcmStaging order."""
        st, ncol = self.staging, self.staging.ncol
        for r in range(self.world):
            off = 0
            for n, h in st.host.items():
                cnt = h.numel()
                src = gcm_all[n][r * ncol:(r + 1) * ncol]
                src = src if isinstance(src, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(src))
                self.inp[r, off:off + cnt].view(h.shape).copy_(src)
                off += cnt

    def _wait(self, word, value):
        import time
        t0 = time.perf_counter()
        while int(self.flags[word]) < value:
            if time.perf_counter() - t0 > self.timeout_s:
                raise RuntimeError("HostExchange: rank %d waited %.0f s for flag %d >= %d" % (self.rank, self.timeout_s, word, value))

    def publish_inputs(self):
        """Owner: the input blocks of the next step are in place."""
        self.step_no += 1
        self.flags[0] = self.step_no

    def fetch_inputs(self, pipe):
        """All ranks, once per step (the owner after fill_inputs): wait for the step's inputs and stage this rank's
        block on its device. Returns the rank's device views (pipe.staging.dev)."""
        if self.rank == self.owner:
            self.publish_inputs()
        else:
            self.step_no += 1
        self._wait(0, self.step_no)
        pipe.staging.dev_buf.copy_(self.inp[self.rank], non_blocking=True)
        return pipe.staging.dev

    def put_tendencies(self, pipe):
        """All ranks, once per step after K3: this rank's tendency block goes to its rows of the shared `out`; the
        owner returns when every rank's block has landed (`out` is complete on the owner only)."""
        lo = self.rank * pipe.ncol
        self.out[lo:lo + pipe.ncol].copy_(pipe.tend, non_blocking=True)
        if pipe.tend.is_cuda:
            torch.cuda.current_stream(pipe.tend.device).synchronize()
        self.flags[1 + self.rank] = self.step_no
        if self.rank == self.owner:
            for r in range(self.world):
                self._wait(1 + r, self.step_no)
        return self.out

    def step(self, pipe, dt=900.0, f_les=1.0, f_gcm=1.0):
        """One sharded host-to-host step (all ranks call it; the owner refreshes the inputs with fill_inputs()
        beforehand when the GCM has moved on). Returns (forcings, out) - `out` is complete on the owner only."""
        self.fetch_inputs(pipe)
        frc = pipe.step_device(dt, f_les, f_gcm)
        return frc, self.put_tendencies(pipe)


class CouplingPipeline(object):
    """State + step of the GPU coupling path for this rank's columns."""

    def __init__(self, cpl, zf, zh, ncol, nlev, dtype=torch.float32, couple_surface=True, layout="kji",
                 ql_thresh=0.0, group=None, gather=True):
        self.cpl = cpl
        dev = cpl.device
        self.zf = torch.as_tensor(np.asarray(zf, dtype=np.float64)).to(dev)
        self.zh = torch.as_tensor(np.asarray(zh, dtype=np.float64)).to(dev)
        self.nk = int(self.zf.shape[0])
        self._zf_top = float(np.asarray(zf, dtype=np.float64)[-1])
        self._tend_live = None
        self.ncol, self.nlev, self.dtype = ncol, nlev, dtype
        self.couple_surface, self.layout, self.ql_thresh = couple_surface, layout, ql_thresh
        self.staging = GcmStaging(ncol, nlev, dtype, dev)
        self.gcm = self.staging.dev
        self.vols = None            # five LES volumes (device-resident LES state)
        self.aux = None             # LES-internal profiles: QL_ice, T, Rhobf, PS ...
        self.slab = None            # last K1 result
        self.tend = torch.zeros((ncol, 7, nlev), dtype=dtype, device=dev)
        self.group = group
        self.world = 1
        self.rank = 0
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(group)
            self.rank = torch.distributed.get_rank(group)
        self.gather = gather and self.world > 1
        # gather mode: True / "nccl" = all_gather_into_tensor after K3; "p2p" = fused gather, K3 stores
        # its block straight into every rank's gather buffer over NVLink (symmetric memory) and a
        # device-side barrier replaces the collective.
        #   "p2p-owner": same, but only the GCM-owning rank (rank 0) receives the blocks - all the path needs.
        self.gather_mode = gather if (self.gather and gather in ("p2p", "p2p-owner")) else "nccl"
        self.symm = None
        self.peer_ptrs = None
        if self.gather and self.gather_mode != "nccl":
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else torch.distributed.group.WORLD
            # two gather buffers used alternately: while a rank still reads step n's buffer (D2H to the
            # GCM), its peers may already store step n+1 into the other one; one barrier per step suffices
            self._p2p = []
            for _ in range(2):
                buf = symm_mem.empty((ncol * self.world, 7, nlev), dtype=dtype, device=dev)
                buf.zero_()
                hdl = symm_mem.rendezvous(buf, grp)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                if self.gather_mode == "p2p-owner":
                    ptrs = [ptrs[0]]                      # rank 0 owns the GCM
                self._p2p.append((buf, hdl, ptrs))
            self._p2p_step = 0
            self.tend_all, self.symm, self.peer_ptrs = self._p2p[0]
        else:
            self.tend_all = torch.zeros((ncol * self.world, 7, nlev), dtype=dtype, device=dev) if self.gather else self.tend
        self.tend_host = torch.empty(self.tend_all.shape, dtype=dtype, pin_memory=True)
        self.k1_events = None       # optional [(start, end)] CUDA events around K1 (bench roofline)
        self._graph = None          # CUDA graph of step_device (capture())
        self._graph_frc = self._graph_args = None
        self._graph_launches = 0

    # LES side --------------------------------------------------------------------------------
    def attach_les(self, vols, aux):
        self.vols = list(vols)
        self.aux = aux

    def les_profiles(self):
        """K1 (get_les_profiles, spcpl.py:747-767)."""
        if self.k1_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.slab = self.cpl.slab_reduce(self.vols, layout=self.layout, ql_thresh=self.ql_thresh,
                                         want_mask=True, out=self.slab)
        if self.k1_events is not None:
            e1.record()
            self.k1_events.append((e0, e1))
        return self.slab

    # one coupled step -------------------------------------------------------------------------
    def forcings(self, dt, factor):
        """K2 (set_les_forcings for all columns, spcpl.py:299-385)."""
        prof = self.slab["prof"] if self.slab is not None else None
        return self.cpl.gcm_to_les(self.gcm, self.zf, self.zh, prof, self.aux["PS"] if prof is not None else None,
                                   dt, factor, self.couple_surface)

    def tendencies(self, frc, dt, factor, conservative=False):
        """K3 (set_gcm_tendencies for all columns, spcpl.py:388-555) + the tendency gather."""
        if self.gather and self.gather_mode != "nccl":
            self.tend_all, self.symm, self.peer_ptrs = self._p2p[self._p2p_step & 1]
            self._p2p_step += 1
            res = self.cpl.les_to_gcm(self.gcm, self.zf, self.zh, self.slab, self.aux, frc["slab_idx"], dt, factor,
                                      conservative=conservative, tend_out=self.tend, peer_ptrs=self.peer_ptrs,
                                      peer_col0=self.rank * self.ncol)
            self.symm.barrier(channel=0)     # every rank's NVLink stores have landed everywhere
            return res
        res = self.cpl.les_to_gcm(self.gcm, self.zf, self.zh, self.slab, self.aux, frc["slab_idx"], dt, factor,
                                  conservative=conservative, tend_out=self.tend)
        if self.gather:
            gather_tendencies(self.tend, self.tend_all, self.group)
        return res

    def step_device(self, dt=900.0, f_les=1.0, f_gcm=1.0):
        """K2 -> K1 -> K3 (+gather) with everything already resident in HBM."""
        frc = self.forcings(dt, f_les)
        self.les_profiles()
        self.tendencies(frc, dt, f_gcm)
        return frc

    # CUDA graph of the device step ---------------------------------------------------------------
    def capture(self, dt=900.0, f_les=1.0, f_gcm=1.0, warmup=2):
        """Records K2 -> K1 -> cloud projection -> K3 of a single-rank step once into a CUDA graph, so that a step is
        one graph launch instead of four calls through the C ABI. Worth it for small column batches, where the
        step is launch-bound; at thousands of columns the host is far ahead of K1 anyway. All buffers of the
        step (forcings, slab means, mask, tendencies) become static: `step_graph()` returns the same tensors
        every time. Sharded runs (tendency gather on) keep the eager path."""
        if self.gather:
            raise RuntimeError("capture() records the single-rank step; with a tendency gather (gather=%r) run it eagerly "
                               "(capturing the collective hung in testing and the fused gather's barrier is not capturable)"
                               % self.gather_mode)
        if self.k1_events is not None:
            raise RuntimeError("K1 event timing must be off during capture")
        cur = torch.cuda.current_stream(self.cpl.device)
        side = torch.cuda.Stream(self.cpl.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up off the capture: lazy kernel attributes, allocator pools
            for _ in range(warmup):
                self.step_device(dt, f_les, f_gcm)
        cur.wait_stream(side)
        l0 = self.cpl.launches
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            frc = self.step_device(dt, f_les, f_gcm)
        self._graph, self._graph_frc, self._graph_launches = graph, frc, self.cpl.launches - l0
        self._graph_args = (dt, f_les, f_gcm)
        return graph

    def step_graph(self):
        """Replays the captured step on the current stream; returns the (static) forcing tensors."""
        self._graph.replay()
        self.cpl.launches += self._graph_launches
        return self._graph_frc

    def first_live_level(self):
        """Host-side bound of where tendencies can be non-zero: the smallest start_index over this rank's columns
        (spcpl.py:494-498: GCM full levels strictly above the LES top get zero), computed from the staged HOST
        profiles with the kernels' own float64 expression (Zgfull - Zghalf[-1]) / grav > zf[-1], minus one level of
        margin. Levels [0, first_live) of every tendency are exactly zero and need not travel."""
        h = self.staging.host
        zf = (h["Zgfull"].double() - h["Zghalf"].double()[:, -1:]) / 9.81
        start = (zf > float(self._zf_top)).sum(dim=1)
        return max(int(start.min()) - 1, 0)

    def step_host(self, dt=900.0, f_les=1.0, f_gcm=1.0, owner=0, compact=False):
        """The step as the host GCM sees it: GCM profiles in pinned host memory in, tendencies in
        pinned host memory out on the rank that owns the GCM. Synchronises before returning.
        compact=True (single rank): only the levels that can be non-zero come back - returns
        (forcings, (tend[:, :, first:], first)) with the block contiguous in pinned memory; `first` comes from
        first_live_level(), evaluated on the host while the GPU runs the step."""
        self.staging.upload()
        if self._graph is not None and self._graph_args == (dt, f_les, f_gcm):
            frc = self.step_graph()
        else:
            frc = self.step_device(dt, f_les, f_gcm)
        if compact and not self.gather:
            first = self.first_live_level()
            nl = self.nlev - first
            if self._tend_live is None:
                self._tend_live = torch.empty(self.tend.numel(), dtype=self.dtype, device=self.cpl.device)
            n = self.ncol * 7 * nl
            dev = self._tend_live[:n].view(self.ncol, 7, nl)
            dev.copy_(self.tend[:, :, first:])                      # strided -> contiguous, on the device
            host = self.tend_host.view(-1)[:n].view(self.ncol, 7, nl)
            host.copy_(dev, non_blocking=True)
            torch.cuda.current_stream(self.cpl.device).synchronize()
            return frc, (host, first)
        if self.rank == owner:
            self.tend_host.copy_(self.tend_all, non_blocking=True)
        torch.cuda.current_stream(self.cpl.device).synchronize()
        return frc, self.tend_host
