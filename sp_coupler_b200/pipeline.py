"""Device-resident batched coupling step for the columns one rank owns.

This is what the `splib.step`-shaped driver (sp_coupler_b200/splib.py) and bench.py run per GCM
time step instead of the reference's two serial Python loops over LES models
(splib/splib.py:317-323 and :330-332):

    H2D  GCM profiles of the step (one packed copy from pinned host memory)   <- gather_gcm_data
    K2   gcm_to_les   forcings on the LES from the previous slab means        <- set_les_forcings
         [ the LES models time-step here; external to the coupling path ]
    K1   slab_reduce  slab means + cloud mask of the LES volumes              <- get_les_profiles
    K3   les_to_gcm   cloud projection + tendencies, packed [ncol][7][nlev];  <- set_gcm_tendencies
                      its epilogue also delivers the block: NVLink stores into the GCM owner's gather buffer
                      (columns sharded over GPUs) or PCIe stores into the host GCM's pinned memory, and a
                      completion flag - no separate collective, copy or host synchronisation

Columns are independent, so ranks own contiguous column blocks and the only exchange is the
tendency gather. The whole step - sharded or not - is three kernel launches with no
host-side barrier, so it is recorded once into a CUDA graph (`capture()`).

Level window. Tendencies of GCM levels above the LES top are zero (spcpl.py:494-533) and the
forcings only need the GCM levels up to the first one above the LES top, so the step gives the same
numbers when the GCM columns are cut off above that level: the kernels simply run on `nlw = nlev - lev0`
levels. `set_levels()` re-views all profile buffers for a window; the host-facing steps use it so that
only live levels cross PCIe (L91: 27 of 91 levels, L137: 38 of 137).
"""
import os
import time

import numpy as np
import torch

from . import _abi
from .constants import grav, surf_vars
from .coupler import GCM_FULL, GCM_HALF, RemoteTargets


def shard_columns(ncol_total, world_size, rank):
    """Contiguous block partition: rank r owns [lo, hi). Remainder columns go to the first ranks."""
    base, rem = divmod(ncol_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_tendencies(tend_local, tend_all, group=None):
    """The path's only exchange as a library collective: all_gather of the packed
    [ncol_local][7][nlev] tendency blocks into [ncol_local*world][7][nlev] so the rank that owns the GCM holds every
    column's tendencies (reference analogue: 7 set_profile_tendency RPCs per column, spcpl.py:535-542).
    NCCL on device tensors, gloo on CPU tensors (tests). The default multi-GPU path does not call this: K3 stores
    its block into the owner's buffer itself (CouplingPipeline, gather="p2p-owner")."""
    torch.distributed.all_gather_into_tensor(tend_all, tend_local, group=group)
    return tend_all


def first_live_level(zgfull, zg_surface, zf_top, margin=1):
    """Host-side bound of where tendencies can be non-zero: the smallest start_index over the given columns
    (spcpl.py:494-498: GCM full levels strictly above the LES top get zero tendencies), computed from the HOST
    profiles with the kernels' own float64 expression (Zgfull - Zghalf[-1]) / grav > zf[-1], minus `margin` levels
    (one level above the LES top is needed as the upper bracket of the GCM->LES interpolation, spcpl.py:224-228).
    zgfull [ncol][nlev], zg_surface [ncol] (= Zghalf[:, -1]); numpy arrays or CPU tensors."""
    zgfull = torch.as_tensor(np.asarray(zgfull) if not isinstance(zgfull, torch.Tensor) else zgfull).double()
    zs = torch.as_tensor(np.asarray(zg_surface) if not isinstance(zg_surface, torch.Tensor) else zg_surface).double()
    zf = (zgfull - zs.reshape(-1, 1)) / grav
    start = (zf > float(zf_top)).sum(dim=1)
    return max(int(start.min()) - int(margin), 0)


def window_columns(gcm, lev0):
    """The GCM columns cut off above level lev0: profile arrays [:, lev0:], surface fields unchanged."""
    if lev0 == 0:
        return gcm
    out = {}
    for n, v in gcm.items():
        out[n] = v[:, lev0:] if (n in GCM_FULL or n in GCM_HALF) else v
    return out


class GcmStaging(object):
    """Packed struct-of-arrays staging of the GCM inputs (gcm_vars + surf_vars, spcpl.py:32-33):
    one pinned host buffer, one device buffer, one async copy per step. Sized for `nlev` levels;
    `set_levels(nlw)` re-packs the views for a window of the lowest nlw levels (contiguous prefix of
    both buffers, so the copy stays a single transfer of `nbytes` bytes)."""

    def __init__(self, ncol, nlev, dtype, device, pin=True):
        self.ncol, self.nlev_max, self.dtype = ncol, nlev, dtype
        total = self.numel_for(ncol, nlev)
        self.host_buf = torch.empty(total, dtype=dtype, pin_memory=pin and torch.cuda.is_available())
        self.dev_buf = torch.empty(total, dtype=dtype, device=device)
        self.esize = self.host_buf.element_size()
        self.set_levels(nlev)

    @staticmethod
    def sizes(ncol, nlev):
        return [(n, (ncol, nlev)) for n in GCM_FULL] + [(n, (ncol, nlev + 1)) for n in GCM_HALF] + \
               [(n, (ncol,)) for n in surf_vars]

    @classmethod
    def numel_for(cls, ncol, nlev):
        return sum(int(np.prod(s)) for _, s in cls.sizes(ncol, nlev))

    @staticmethod
    def views(buf, ncol, nlev):
        out, off = {}, 0
        for n, s in GcmStaging.sizes(ncol, nlev):
            cnt = int(np.prod(s))
            out[n] = buf[off:off + cnt].view(*s)
            off += cnt
        return out

    def set_levels(self, nlw):
        if not 2 <= nlw <= self.nlev_max:
            raise ValueError("level window %d outside [2, %d]" % (nlw, self.nlev_max))
        self.nlev = nlw
        self.numel = self.numel_for(self.ncol, nlw)
        self.nbytes = self.numel * self.esize
        self.host = self.views(self.host_buf, self.ncol, nlw)
        self.dev = self.views(self.dev_buf, self.ncol, nlw)

    def fill_host(self, gcm):
        """Copy a dict of numpy / CPU-tensor arrays (already cut to the current level window) into the pinned buffer
        (what the host GCM does)."""
        for n, h in self.host.items():
            if n in gcm:
                h.copy_(torch.as_tensor(np.ascontiguousarray(gcm[n])) if not isinstance(gcm[n], torch.Tensor) else gcm[n])

    def upload(self):
        self.dev_buf[:self.numel].copy_(self.host_buf[:self.numel], non_blocking=True)
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


def _spin_until(read, value, timeout_s, what):
    """Poll `read() >= value`. The wait is one GPU step away (tens of microseconds to a few milliseconds), and a timed
    sleep would add its wake-up latency (50-100 us measured) to every step, so: a short pure spin, then spinning with
    sched_yield() (returns at once when nothing else wants the core, gives it up when something does), and only after
    50 ms - a peer that is really late - sleeps of 0.2 ms so that a long wait does not burn a core. x86-TSO makes a
    plain load sufficient for a word written with a release store by the GPU or by another process."""
    n = 0
    t0 = None
    while True:
        if read() >= value:
            return
        n += 1
        if n <= 2000:
            continue
        if t0 is None:
            t0 = time.perf_counter()
        if n % 64:
            os.sched_yield()
            continue
        el = time.perf_counter() - t0
        if el > timeout_s:
            raise RuntimeError("%s: waited %.0f s" % (what, el))
        if el > 0.05:
            time.sleep(2e-4)


class HostExchange(object):
    """Host side of a sharded step when the GCM lives in HOST memory of one process (OpenIFS does): one pinned
    host buffer shared by all ranks of the node (a /dev/shm mapping that every rank maps for its GPU).

        GCM owner   writes every rank's packed input block (cut to the live level window) into `inp[r]`,
                    then publishes (step number, lev0) in the header
        every rank  waits for it, copies ITS block host->device (one copy over its own PCIe link) and replays the
                    step; K3 stores ITS tendency block straight into the shared buffer's `out[r*ncol:(r+1)*ncol]` and
                    then sets the rank's completion flag there (zero-copy PCIe stores from the kernel's epilogue)
        GCM owner   polls the flags of all ranks: `out` holds [world*ncol][7][nlw] - no device gather, no collective,
                    no device->host copy call and no stream synchronisation on the critical path

    Every rank owns the same number of columns (checked collectively). Reference analogue: the master gathers every
    profile over its own channels and sends every tendency back (spcpl.py:55-86, 535-542).
    Memory ordering: header words are written by the owner process after the input blocks (x86-TSO keeps the order;
    readers poll the step word first); GPU flags are release stores at system scope after a system-wide fence.
    Works on CPU tensors too (no mapping; `host_step` is then driven by the gloo tests with a host-side stand-in)."""

    HEADER_WORDS = 8     # int64: [0] step number of the inputs, [1] lev0 of their level window

    def __init__(self, pipe, world, rank, owner=0, group=None, register=True, tag="x", timeout_s=60.0, window=True,
                 direct=True):
        """direct=True: K3 stores the rank's block into the shared buffer and raises the flag itself (no host call
        after the launch). direct=False: the rank copies its block device->host with the copy engine, synchronises its
        stream and sets the flag from the host (faster for blocks of many MB: the copy engine moves ~54 GB/s, stores
        issued by the SMs ~33 GB/s)."""
        self.pipe, self.world, self.rank, self.owner, self.group = pipe, world, rank, owner, group
        self.timeout_s, self.window, self.direct = timeout_s, window, direct
        staging = pipe.staging
        ncol, nlev, dtype = staging.ncol, staging.nlev_max, staging.dtype
        self.ncol, self.nlev = ncol, nlev
        esize = torch.empty((), dtype=dtype).element_size()
        self.esize = esize
        self.per_rank_in = GcmStaging.numel_for(ncol, nlev)
        nin = world * self.per_rank_in * esize
        nout = world * ncol * 7 * nlev * esize
        al = lambda n: (n + 4095) // 4096 * 4096
        self._off_flags = al(self.HEADER_WORDS * 8)
        self._off_out = self._off_flags + al(_abi.SYNC_WORDS * 4)
        self._off_in = self._off_out + al(nout)
        self.nbytes = self._off_in + al(nin)
        # every phase ends in an all_reduce(MIN) of "it worked here", so that all ranks either get the buffer or raise
        # together - a rank that fails alone must never leave the others waiting in a collective
        dev = "cuda" if torch.distributed.get_backend(group) == "nccl" else "cpu"
        shape = torch.tensor([ncol, -ncol, nlev, -nlev], device=dev)
        torch.distributed.all_reduce(shape, op=torch.distributed.ReduceOp.MIN, group=group)
        if int(shape[0]) != -int(shape[1]) or int(shape[2]) != -int(shape[3]):
            raise ValueError("HostExchange needs the same number of columns and levels on every rank "
                             "(columns %d..%d, levels %d..%d)" % (int(shape[0]), -int(shape[1]), int(shape[2]), -int(shape[3])))
        err, path = None, None
        try:
            if rank == owner:       # private, unpredictable name; O_EXCL | O_NOFOLLOW semantics of mkstemp, mode 0600
                import tempfile
                fd, path = tempfile.mkstemp(prefix="spcpl_b200_%s_" % tag, dir="/dev/shm")
                os.ftruncate(fd, self.nbytes)
                os.close(fd)
        except Exception as e:          # noqa: BLE001
            err = e
        ok = self._agree(err is None, group)
        names = [path]
        if ok:
            torch.distributed.broadcast_object_list(names, src=owner, group=group)
            path = names[0]
        self.raw, self.registered, self.dev_base = None, False, None
        if ok:
            try:
                self.raw = torch.from_file(path, shared=True, size=self.nbytes, dtype=torch.uint8)
                # first touch: this rank's input and output blocks are allocated on the NUMA node it runs on
                self.raw[self._off_in + rank * self.per_rank_in * esize:self._off_in + (rank + 1) * self.per_rank_in * esize].zero_()
                o0 = self._off_out + rank * ncol * 7 * nlev * esize
                self.raw[o0:o0 + ncol * 7 * nlev * esize].zero_()
                if register and pipe.cpl is not None:
                    self.dev_base = pipe.cpl.host_register(self.raw)
                    self.registered = True
            except Exception as e:      # noqa: BLE001
                err = e
            ok = self._agree(err is None, group)
        if rank == owner and path is not None:
            try:
                os.unlink(path)        # the mapping stays alive in every process; nothing is left behind
            except OSError:
                pass
        if not ok:
            self.close()
            raise RuntimeError("HostExchange: shared pinned host buffer unavailable on at least one rank (%s)" % (err,))
        self.header = self.raw[:self.HEADER_WORDS * 8].view(torch.int64)
        self.flags = self.raw[self._off_flags:self._off_flags + _abi.SYNC_WORDS * 4].view(torch.int32)
        self._out_flat = self.raw[self._off_out:self._off_out + nout].view(dtype)
        self._in_flat = self.raw[self._off_in:self._off_in + nin].view(dtype)
        self.inp = self._in_flat.view(world, self.per_rank_in)
        if rank == owner:
            self.header.zero_()
            self.flags.zero_()
        self._hdr_np, self._flags_np = self.header.numpy(), self.flags.numpy()    # plain loads for the polling loops
        self.lev0 = 0
        self.step_no = 0
        pipe.upload_source(self.inp[rank])      # the step's H2D copy reads this rank's block of the shared buffer
        if self.registered and direct:      # K3 of this rank stores into the shared buffer and signals its flag there
            pipe.bind_host_output(self.dev_base + self._off_out, self.dev_base + self._off_flags, col0=rank * ncol, slot=rank)
        torch.distributed.barrier(group=group)

    @staticmethod
    def _agree(ok, group):
        dev = "cuda" if torch.distributed.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor([1 if ok else 0], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=group)
        return bool(int(t.item()))

    def close(self):
        """Unmap the shared buffer from this rank's GPU (after its queued work has finished). The pipeline must not be
        stepped through this exchange afterwards: its K3 targets and upload source point into the buffer."""
        if self.registered:
            torch.cuda.synchronize()
            self.pipe.cpl.host_unregister(self.raw)
            self.registered = False

    def __del__(self):
        try:
            self.close()
        except Exception:           # noqa: BLE001 - interpreter shutdown: the driver unmaps the memory itself
            pass

    def out(self, nlw=None):
        """[world*ncol][7][nlw] view of the shared output region for the current (or given) level window."""
        nlw = self.nlev - self.lev0 if nlw is None else nlw
        return self._out_flat[:self.world * self.ncol * 7 * nlw].view(self.world * self.ncol, 7, nlw)

    def fill_inputs(self, gcm_all, zf_top=None):
        """Owner only. gcm_all: dict of [world*ncol, ...] host arrays (all nlev levels) in global column order - what
        gather_gcm_data fetched from the host GCM. Chooses the level window of the step from the heights of ALL
        columns and packs every rank's block, cut to the window, in GcmStaging order."""
        ncol = self.ncol
        lev0 = 0
        if self.window:
            top = self.pipe._zf_top if zf_top is None else zf_top
            lev0 = first_live_level(gcm_all["Zgfull"], np.asarray(gcm_all["Zghalf"])[:, -1], top)
        self._next_lev0 = lev0
        nlw = self.nlev - lev0
        per = GcmStaging.numel_for(ncol, nlw)
        for r in range(self.world):
            views = GcmStaging.views(self.inp[r, :per], ncol, nlw)
            for n, h in views.items():
                src = gcm_all[n][r * ncol:(r + 1) * ncol]
                if n in GCM_FULL or n in GCM_HALF:
                    src = src[:, lev0:]
                src = src if isinstance(src, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(src))
                h.copy_(src)

    def publish_inputs(self):
        """Owner: the input blocks of the next step are in place."""
        self.step_no += 1
        self.header[1] = getattr(self, "_next_lev0", 0)
        self.header[0] = self.step_no

    def fetch_inputs(self, stage=True):
        """All ranks, once per step (the owner after fill_inputs): wait for the step's inputs, adopt their level
        window and (stage=True) copy this rank's block to its device. Returns the rank's device views
        (pipe.staging.dev)."""
        pipe = self.pipe
        if self.rank == self.owner:
            self.publish_inputs()
        else:
            self.step_no += 1
        hdr = self._hdr_np
        _spin_until(lambda: int(hdr[0]), self.step_no, self.timeout_s, "HostExchange rank %d: inputs of step %d" % (self.rank, self.step_no))
        self.lev0 = int(hdr[1])
        nlw = self.nlev - self.lev0
        if pipe.nlw != nlw:
            pipe.set_levels(nlw)
        if stage:
            pipe._upload()
        return pipe.staging.dev

    def wait_tendencies(self):
        """After the step has been launched. The owner returns when every rank's K3 has signalled that its block is in
        the shared buffer (`out()` is complete on the owner only); the other ranks return at once. Without a mapped
        buffer (CPU tests) every rank copies its block and sets its flag from the host."""
        pipe = self.pipe
        if not (self.registered and self.direct):
            lo = self.rank * self.ncol
            self.out()[lo:lo + self.ncol].copy_(pipe.tend, non_blocking=True)
            if pipe.tend.is_cuda:
                torch.cuda.current_stream(pipe.tend.device).synchronize()
            pipe.epoch += 1
            self._flags_np[_abi.SYNC_FLAG0 + self.rank] = pipe.epoch
        if self.rank == self.owner:
            fl, want = self._flags_np, pipe.epoch
            for r in range(self.world):
                _spin_until(lambda: int(fl[_abi.SYNC_FLAG0 + r]), want, self.timeout_s,
                            "HostExchange: completion flag of rank %d (step %d)" % (r, self.step_no))
        return self.out()

    def step(self, dt=900.0, f_les=1.0, f_gcm=1.0):
        """One sharded host-to-host step (all ranks call it; the owner refreshes the inputs with fill_inputs()
        beforehand when the GCM has moved on). Returns (forcings, out, lev0): `out` [world*ncol][7][nlev-lev0] is
        complete on the owner only; tendencies of the levels above lev0 are zero."""
        self.fetch_inputs(stage=False)
        frc = self.pipe.step(dt, f_les, f_gcm, upload=True)      # H2D of this rank's block + K2 -> K1 -> K3: one graph launch
        return frc, self.wait_tendencies(), self.lev0


class CouplingPipeline(object):
    """State + step of the GPU coupling path for this rank's columns.

    gather: how the packed tendency block reaches the rank that owns the GCM when columns are sharded
      False / None   stays on this rank (single GPU, or the host exchange delivers it)
      "p2p-owner"    K3 stores it into the owner's gather buffer over NVLink (all the path needs); default in bench.py
      "p2p"          K3 stores it into every rank's gather buffer (an all-gather)
      "nccl" / True  all_gather_into_tensor after K3 (the library collective; not graph-captured)
    With a p2p mode the launch ends in K3's own device-side barrier (flags in symmetric memory), so the sharded step
    needs no NCCL call and no host synchronisation."""

    def __init__(self, cpl, zf, zh, ncol, nlev, dtype=torch.float32, couple_surface=True, layout="kji",
                 ql_thresh=0.0, group=None, gather=True, owner=0):
        self.cpl = cpl
        dev = cpl.device
        self.zf = torch.as_tensor(np.asarray(zf, dtype=np.float64)).to(dev)
        self.zh = torch.as_tensor(np.asarray(zh, dtype=np.float64)).to(dev)
        self.nk = int(self.zf.shape[0])
        self._zf_top = float(np.asarray(zf, dtype=np.float64)[-1])
        self.ncol, self.nlev, self.dtype = ncol, nlev, dtype
        self.couple_surface, self.layout, self.ql_thresh = couple_surface, layout, ql_thresh
        self.staging = GcmStaging(ncol, nlev, dtype, dev)
        self.vols = None            # five LES volumes (device-resident LES state)
        self.aux = None             # LES-internal profiles: QL_ice, T, Rhobf, PS ...
        self.slab = None            # last K1 result
        self.group = group
        self.world, self.rank, self.owner = 1, 0, owner
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(group)
            self.rank = torch.distributed.get_rank(group)
        self.gather = bool(gather) and self.world > 1
        self.gather_mode = (gather if gather in ("p2p", "p2p-owner") else "nccl") if self.gather else None
        self._tend_buf = torch.zeros(ncol * 7 * nlev, dtype=dtype, device=dev)
        self._host_buf = None       # pinned mirror of the (gathered) block, allocated on first use
        self.remote = None          # RemoteTargets of K3 (peer gather buffers or pinned host memory)
        self.epoch = 0              # host mirror of the sync epoch: K3 launches with the completion protocol so far
        self._p2p = None
        self._all_buf = None
        self._host_flags = None
        if self.gather and self.gather_mode != "nccl":
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else torch.distributed.group.WORLD
            # two gather buffers used alternately (selected by the epoch inside K3): while a rank still reads step n's
            # buffer (D2H to the GCM), its peers may already store step n+1 into the other one
            bufs, sets = [], []
            for _ in range(2):
                buf = symm_mem.empty(ncol * self.world * 7 * nlev, dtype=dtype, device=dev)
                buf.zero_()
                hdl = symm_mem.rendezvous(buf, grp)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                sets.append([ptrs[owner]] if self.gather_mode == "p2p-owner" else ptrs)
                bufs.append((buf, hdl))
            sync = symm_mem.empty(_abi.SYNC_WORDS, dtype=torch.int32, device=dev)
            sync.zero_()
            shdl = symm_mem.rendezvous(sync, grp)
            self._p2p = (bufs, sync, shdl)
            self.remote = RemoteTargets(sets, col0=self.rank * ncol, sync=sync, signal=[int(p) for p in shdl.buffer_ptrs],
                                        slot=self.rank, n_wait=self.world)
            torch.cuda.synchronize(dev)
            torch.distributed.barrier(group=group)      # every sync block is zero before anyone's first K3
        elif self.gather:
            self._all_buf = torch.zeros(ncol * self.world * 7 * nlev, dtype=dtype, device=dev)
        self.k1_events = None       # optional [(start, end)] CUDA events around K1 (bench roofline)
        self._graphs = {}           # (nlw, dt, f_les, f_gcm, with upload) -> (graph, forcings, launches)
        self._upload_src = None     # host side of the step's H2D copy when it is not the own staging buffer
        self._out_cache = {}        # nlw -> (K2 outputs, K3 outputs): written in place every step
        self.set_levels(nlev)

    # level window ------------------------------------------------------------------------------
    def set_levels(self, nlw):
        """Run the step on the lowest `nlw` GCM levels (module docstring, "Level window"). Re-views the staging
        buffers, the packed tendency block and the gather / host targets; the buffers themselves stay."""
        self.staging.set_levels(nlw)
        self.nlw = nlw
        self.gcm = self.staging.dev
        n = self.ncol * 7 * nlw
        self.tend = self._tend_buf[:n].view(self.ncol, 7, nlw)
        return self

    @property
    def tend_all(self):
        """The gathered block [ncol*world][7][nlw] of the LAST completed step (valid on the GCM owner; on every rank
        with gather="p2p"/"nccl"); the local block when nothing is gathered."""
        n = self.ncol * self.world * 7 * self.nlw
        if self._p2p is not None:
            buf = self._p2p[0][(self.epoch - 1) & 1][0] if self.epoch > 0 else self._p2p[0][0][0]
            return buf[:n].view(self.ncol * self.world, 7, self.nlw)
        if self._all_buf is not None:
            return self._all_buf[:n].view(self.ncol * self.world, 7, self.nlw)
        return self.tend

    @property
    def tend_host(self):
        """Pinned host mirror of tend_all for the current level window."""
        total = self.ncol * self.world * 7 * self.nlev
        if self._host_buf is None:
            self._host_buf = torch.zeros(total, dtype=self.dtype, pin_memory=True)
        n = self.ncol * self.world * 7 * self.nlw
        return self._host_buf[:n].view(self.ncol * self.world, 7, self.nlw)

    def bind_host_output(self, out_ptr=None, flags_ptr=None, col0=0, slot=0):
        """Make K3 deliver this rank's tendency block to HOST memory itself: zero-copy stores into a pinned buffer
        [..][7][nlw] at column offset col0 and a completion flag next to it (RemoteTargets). Without arguments the
        pipeline's own pinned mirror (`tend_host`) and a private flag block are used (single-GPU host step)."""
        if self.gather:
            raise RuntimeError("bind_host_output is for pipelines without a device gather (gather=False)")
        if out_ptr is None:
            self.tend_host                                  # allocate
            self._host_flags = torch.zeros(_abi.SYNC_WORDS, dtype=torch.int32, pin_memory=True)
            self._host_flags_np = self._host_flags.numpy()
            out_ptr = self.cpl.host_device_pointer(self._host_buf)
            flags_ptr = self.cpl.host_device_pointer(self._host_flags)
        self._sync_dev = torch.zeros(_abi.SYNC_WORDS, dtype=torch.int32, device=self.cpl.device)
        self.remote = RemoteTargets([[out_ptr]], col0=col0, sync=self._sync_dev, signal=[flags_ptr], slot=slot, n_wait=0)
        self.epoch = 0
        self._graphs.clear()
        torch.cuda.synchronize(self.cpl.device)
        return self

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
    def _outs(self):
        return self._out_cache.setdefault(self.nlw, [None, None])

    def forcings(self, dt, factor):
        """K2 (set_les_forcings for all columns, spcpl.py:299-385)."""
        prof = self.slab["prof"] if self.slab is not None else None
        outs = self._outs()
        outs[0] = self.cpl.gcm_to_les(self.gcm, self.zf, self.zh, prof, self.aux["PS"] if prof is not None else None,
                                      dt, factor, self.couple_surface, out=outs[0])
        return outs[0]

    def tendencies(self, frc, dt, factor, conservative=False):
        """K3 (set_gcm_tendencies for all columns, spcpl.py:388-555), which also delivers the block (class docstring)."""
        outs = self._outs()
        outs[1] = self.cpl.les_to_gcm(self.gcm, self.zf, self.zh, self.slab, self.aux, frc["slab_idx"], dt, factor,
                                      conservative=conservative, tend_out=self.tend, remote=self.remote, out=outs[1])
        if self.remote is not None and self.remote.sync is not None and not torch.cuda.is_current_stream_capturing():
            self.epoch += 1
        if self.gather and self.gather_mode == "nccl":
            gather_tendencies(self.tend, self.tend_all, self.group)
        return outs[1]

    def step_device(self, dt=900.0, f_les=1.0, f_gcm=1.0):
        """K2 -> K1 -> K3 (incl. the tendency delivery) with everything already resident in HBM."""
        frc = self.forcings(dt, f_les)
        self.les_profiles()
        self.tendencies(frc, dt, f_gcm)
        return frc

    def step(self, dt=900.0, f_les=1.0, f_gcm=1.0, upload=False):
        """The device step: replayed from its CUDA graph when one has been captured for the current level window
        and factors, else launched eagerly. upload=True: the H2D copy of the staged GCM columns is part of the step
        (recorded into the graph by capture(upload=True), else issued here)."""
        g = None
        if upload:
            g = self._graphs.get((self.nlw, dt, f_les, f_gcm, True))
            if g is None:
                self._upload()
        if g is None:
            g = self._graphs.get((self.nlw, dt, f_les, f_gcm, False))
        if g is None:
            return self.step_device(dt, f_les, f_gcm)
        graph, frc, launches = g
        graph.replay()
        self.cpl.launches += launches
        if self.remote is not None and self.remote.sync is not None:
            self.epoch += 1
        return frc

    def _upload(self):
        """H2D of the staged level window from the step's host source (the pinned staging buffer, or the block of a
        shared exchange buffer set with upload_source)."""
        n = self.staging.numel
        src = self._upload_src if self._upload_src is not None else self.staging.host_buf
        self.staging.dev_buf[:n].copy_(src[:n], non_blocking=True)

    def upload_source(self, src):
        """Use `src` (a pinned / registered CPU tensor laid out like the staging buffer) as the host side of the step's
        H2D copy instead of the pipeline's own pinned staging buffer."""
        self._upload_src = src
        for k in [k for k in self._graphs if k[4]]:
            del self._graphs[k]

    # CUDA graph of the device step ---------------------------------------------------------------
    def capture(self, dt=900.0, f_les=1.0, f_gcm=1.0, warmup=2, upload=False):
        """Records K2 -> K1 -> K3 once into a CUDA graph, so that a step is one graph launch instead of three calls
        through the C ABI. Sharded steps are captured too: the gather and its barrier are inside K3. All ranks must
        call it together (the warm-up steps run the device barrier). All buffers of the step are static: `step()`
        returns the same forcing tensors every time. Only the NCCL gather mode keeps the eager path.
        upload=True records the H2D copy of the staged level window (from the upload source) as the graph's first
        node: the host-facing step is then ONE launch."""
        if self.gather and self.gather_mode == "nccl":
            raise RuntimeError("capture() needs the fused gather (gather='p2p' / 'p2p-owner'): the NCCL collective is "
                               "not recorded into the step graph")
        if self.k1_events is not None:
            raise RuntimeError("K1 event timing must be off during capture")
        cur = torch.cuda.current_stream(self.cpl.device)
        side = torch.cuda.Stream(self.cpl.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up off the capture: output buffers exist, allocator pools are warm
            for _ in range(max(warmup, 1)):
                if upload:
                    self._upload()
                self.step_device(dt, f_les, f_gcm)
        cur.wait_stream(side)
        l0 = self.cpl.launches
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            if upload:
                self._upload()
            frc = self.step_device(dt, f_les, f_gcm)
        self._graphs[(self.nlw, dt, f_les, f_gcm, bool(upload))] = (graph, frc, self.cpl.launches - l0)
        self.cpl.launches = l0
        return graph

    def step_graph(self, dt=900.0, f_les=1.0, f_gcm=1.0):
        """Replays the captured step on the current stream; returns the (static) forcing tensors."""
        if (self.nlw, dt, f_les, f_gcm, False) not in self._graphs:
            raise RuntimeError("no graph captured for this level window / factors")
        return self.step(dt, f_les, f_gcm)

    def sync_error(self):
        """Non-zero when a K3 launch gave up waiting for a peer's completion flag (20 s); reads one device word."""
        sync = self.remote.sync if self.remote is not None else None
        return 0 if sync is None else int(sync[_abi.SYNC_ERROR].item())

    # host-facing step --------------------------------------------------------------------------
    def first_live_level(self, gcm_host=None):
        """lev0 of the level window for the given host columns (default: the staged ones, full levels)."""
        h = self.staging.host if gcm_host is None else gcm_host
        return first_live_level(h["Zgfull"], torch.as_tensor(np.asarray(h["Zghalf"]))[:, -1], self._zf_top)

    def stage_host(self, gcm_host, window=True, lev0=None):
        """The host GCM's part of gather_gcm_data: pick the level window of these columns, pack them (cut to the
        window) into the pinned staging buffer. Returns lev0. Not per step unless the GCM state changed.
        When the columns are sharded every rank must run the same window (the gathered block is [..][7][nlw] on the
        owner): pass the job-wide lev0 = min over ranks of first_live_level()."""
        if lev0 is None:
            lev0 = self.first_live_level(gcm_host) if window else 0
        nlw = self.nlev - lev0
        if nlw != self.nlw:
            self.set_levels(nlw)
        self.staging.fill_host(window_columns(gcm_host, lev0))
        self.lev0 = lev0
        return lev0

    def step_host(self, dt=900.0, f_les=1.0, f_gcm=1.0, owner=0):
        """The step as the host GCM sees it: GCM profiles in pinned host memory in (one H2D copy of the staged level
        window), tendencies in pinned host memory out on the rank that owns the GCM. Returns (forcings, tend_host)
        with tend_host [ncol*world][7][nlw] for the current window (levels above it are zero).
        After bind_host_output() K3 writes tend_host itself and the host only polls the completion flag; otherwise the
        (gathered) device block is copied back and the stream synchronised."""
        frc = self.step(dt, f_les, f_gcm, upload=True)
        if self._host_flags is not None:
            fl, want = self._host_flags_np, self.epoch
            _spin_until(lambda: int(fl[_abi.SYNC_FLAG0]), want, 60.0, "step_host: K3 completion flag")
            return frc, self.tend_host
        if self.rank == owner:
            self.tend_host.copy_(self.tend_all, non_blocking=True)
        torch.cuda.current_stream(self.cpl.device).synchronize()
        return frc, self.tend_host
