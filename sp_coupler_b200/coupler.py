"""Batched host interface of the coupling kernels over torch device tensors.

`Coupler` owns one C-ABI handle per device and exposes the path's operations over `[ncol, ...]`
tensors. torch is plumbing only (device memory, streams); every number is produced by the
hand-written sm_100a kernels behind include/spcpl_b200.h. There is no CPU path: CPU tensors are
rejected, and a missing library raises at first use.
"""
import ctypes as C

import torch

from . import _abi
from .constants import LES_FIELDS, TENDENCIES, surf_vars

_DT = {torch.float32: _abi.SPC_F32, torch.float64: _abi.SPC_F64}
_LAYOUT = {"kji": _abi.LAYOUT_KJI, "ijk": _abi.LAYOUT_IJK, 0: 0, 1: 1}
GCM_FULL = ("U", "V", "T", "SH", "QL", "QI", "Pfull", "A", "Zgfull")
GCM_HALF = ("Phalf", "Zghalf")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class RemoteTargets(object):
    """Where K3 sends this rank's tendency block besides its local buffer, and how the launch reports completion
    (struct spc_gcm_tend: tend_peers / sync / signal, include/spcpl_b200.h). Built once, reused every step; holds the
    ctypes arrays alive.

      targets    list of buffer sets, each a list of device-visible addresses of [ncol_total][7][nlev] buffers
                 (peer GPU memory over NVLink, or pinned host memory); 1 or 2 sets, used alternately
      col0       column offset of this rank's block inside a target
      sync       this rank's sync block (int32 device tensor of _abi.SYNC_WORDS words, zeroed) or None
      signal     addresses of the sync blocks whose flag [slot] the launch sets when all its stores are visible
      n_wait     the launch then waits for flags [0, n_wait) of its own block (device-side barrier)"""

    def __init__(self, targets=(), col0=0, sync=None, signal=(), slot=0, n_wait=0):
        sets = [list(t) for t in targets if len(t)]
        if len(sets) not in (0, 1, 2) or any(len(t) != len(sets[0]) for t in sets):
            raise ValueError("targets: one or two buffer sets of equal length")
        self.n_bufs = max(len(sets), 1)
        self.n_peers = len(sets[0]) if sets else 0
        if self.n_peers > _abi.MAX_PEERS or len(signal) > _abi.MAX_PEERS + 1:
            raise ValueError("at most %d targets" % _abi.MAX_PEERS)
        if len(sets) == 2 and sync is None:
            raise ValueError("two buffer sets alternate by the sync epoch: pass a sync block")
        flat = [int(p) for t in sets for p in t]
        self._parr = (C.c_void_p * max(len(flat), 1))(*flat)
        self._sarr = (C.c_void_p * max(len(signal), 1))(*[int(p) for p in signal])
        self.col0, self.sync, self.n_signal, self.slot, self.n_wait = int(col0), sync, len(signal), int(slot), int(n_wait)
        if sync is not None and (sync.dtype != torch.int32 or sync.numel() < _abi.SYNC_WORDS or not sync.is_contiguous()):
            raise ValueError("sync must be a contiguous int32 tensor of %d words" % _abi.SYNC_WORDS)

    def fill(self, o):
        if self.n_peers:
            o.tend_peers = C.cast(self._parr, C.c_void_p)
            o.n_peers, o.peer_col0 = self.n_peers, self.col0
        o.n_bufs = self.n_bufs
        if self.sync is not None:
            o.sync = self.sync.data_ptr()
            o.signal = C.cast(self._sarr, C.c_void_p)
            o.n_signal, o.sync_slot, o.n_wait = self.n_signal, self.slot, self.n_wait


class Coupler(object):
    """Per-device entry point. All methods are asynchronous on torch's current stream."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("sp_coupler_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _abi.lib()
        h = C.c_void_p()
        _abi.check(self._lib.spc_create(C.byref(h), self.device.index), "spc_create")
        self._h = h
        self.launches = 0   # kernels launched through this handle (bench.py's gpu_launches)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.spc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ host exchange memory
    def host_register(self, tensor):
        """Pin the host memory of a CPU tensor (e.g. a /dev/shm mapping shared by the ranks of a node) and map it for
        this device; returns the device-visible address K3 may store into (RemoteTargets)."""
        dp = C.c_void_p()
        _abi.check(self._lib.spc_host_register(self._h, C.c_void_p(tensor.data_ptr()), tensor.numel() * tensor.element_size(),
                                               C.byref(dp)), "spc_host_register")
        return int(dp.value)

    def host_unregister(self, tensor):
        _abi.check(self._lib.spc_host_unregister(self._h, C.c_void_p(tensor.data_ptr())), "spc_host_unregister")

    def host_device_pointer(self, tensor):
        """Device-visible address of an already pinned CPU tensor (torch pin_memory)."""
        dp = C.c_void_p()
        _abi.check(self._lib.spc_host_device_pointer(self._h, C.c_void_p(tensor.data_ptr()), C.byref(dp)),
                   "spc_host_device_pointer")
        return int(dp.value)

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t, name, dtype=None, shape=None):
        if not isinstance(t, torch.Tensor):
            raise TypeError("%s must be a torch tensor" % name)
        if t.device != self.device:
            raise ValueError("%s is on %s, expected %s" % (name, t.device, self.device))
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous" % name)
        if dtype is not None and t.dtype != dtype:
            raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
        return t

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _gcm_struct(self, gcm, couple_surface):
        T = gcm["T"]
        dtype = T.dtype
        if dtype not in _DT:
            raise TypeError("GCM profiles must be float32 or float64")
        ncol, nlev = T.shape
        s = _abi.GcmCols()
        s.ncol, s.nlev, s.dtype = ncol, nlev, _DT[dtype]
        for n in GCM_FULL:
            setattr(s, n, self._chk(gcm[n], n, dtype, (ncol, nlev)).data_ptr())
        for n in GCM_HALF:
            setattr(s, n, self._chk(gcm[n], n, dtype, (ncol, nlev + 1)).data_ptr())
        if couple_surface:
            for n in surf_vars:
                if n == "TLflux" and n not in gcm:
                    continue
                setattr(s, n, self._chk(gcm[n], n, dtype, (ncol,)).data_ptr())
        return s, ncol, nlev, dtype

    # ------------------------------------------------------------------ K1
    def mask_words_per_column(self, dtype, layout, nx, ny, nk):
        return int(self._lib.spc_mask_words_per_column(_DT[dtype], _LAYOUT[layout], nx, ny, nk))

    def slab_reduce(self, vols, layout="kji", ql_thresh=0.0, want_cnt=True, want_mask=True, out=None):
        """Slab means + cloud count of the five LES volumes (spcpl.py:747-759,765).

        vols: dict or sequence (THL,QT,QL,U,V) of [ncol,nk,ny,nx] ('kji') or [ncol,nx,ny,nk] ('ijk').
        Returns dict(prof=[5,ncol,nk] f64, cnt=[ncol,nk] i32 | None, mask | None, nx, ny, dtype, layout)."""
        v = [vols[f] for f in LES_FIELDS] if isinstance(vols, dict) else list(vols)
        lay = _LAYOUT[layout]
        dtype = v[0].dtype
        if lay == _abi.LAYOUT_KJI:
            ncol, nk, ny, nx = v[0].shape
        else:
            ncol, nx, ny, nk = v[0].shape
        for i, t in enumerate(v):
            self._chk(t, "vol[%s]" % LES_FIELDS[i], dtype, v[0].shape)
        # buffers of a previous result are reused only if they still fit this call (shape, dtype, device): a stale
        # mask or count buffer of another shape / layout would be written out of bounds
        out = {} if out is None else out

        def reuse(name, shape, dt_):
            t = out.get(name)
            if isinstance(t, torch.Tensor) and tuple(t.shape) == tuple(shape) and t.dtype == dt_ and \
                    t.device == self.device and t.is_contiguous():
                return t
            return self._empty(shape, dt_)

        prof = reuse("prof", (5, ncol, nk), torch.float64)
        cnt = reuse("cnt", (ncol, nk), torch.int32) if want_cnt else None
        mask = None
        if want_mask:
            mw = self.mask_words_per_column(dtype, lay, nx, ny, nk)
            if mw == 0:
                raise RuntimeError("no cloud-mask format for layout %r; pass want_mask=False" % (layout,))
            mask = reuse("mask", (ncol, mw), torch.int32)
        arr = (C.c_void_p * 5)(*[t.data_ptr() for t in v])
        _abi.check(self._lib.spc_slab_reduce(self._h, arr, _DT[dtype], lay, ncol, nx, ny, nk, float(ql_thresh),
                                             _ptr(prof), _ptr(cnt), _ptr(mask), self._stream()), "spc_slab_reduce")
        self.launches += 1
        return dict(prof=prof, cnt=cnt, mask=mask, nx=nx, ny=ny, dtype=dtype, layout=lay)

    # ------------------------------------------------------------------ K2
    def _alloc_into(self, res, o, out):
        """alloc(name, shape, dtype): take a fitting tensor from `out` (validated) or allocate; record it in the
        result dict and in the ctypes struct."""
        def alloc(name, shape, dt_):
            t = out.get(name) if out else None
            if t is None:
                t = self._empty(shape, dt_)
            else:
                self._chk(t, "out[%r]" % name, dt_, shape)
            res[name] = t
            setattr(o, name, t.data_ptr())
        return alloc

    def gcm_to_les(self, gcm, zf, zh=None, les_prof=None, ps_les=None, dt=900.0, factor=1.0,
                   couple_surface=True, diagnostics=False, want_state=False, want_bracket=False, out=None):
        """convert_profiles + set_les_forcings arithmetic + convert_surface_fluxes for all columns
        (spcpl.py:171-246, 299-385, 136-167). Returns a dict of output tensors; `out` (a previous result of the same
        call shape) is written in place instead of allocating."""
        s, ncol, nlev, dtype = self._gcm_struct(gcm, couple_surface)
        nk = zf.shape[0]
        self._chk(zf, "zf", torch.float64, (nk,))
        if zh is not None:
            self._chk(zh, "zh", torch.float64, (nk,))
        if les_prof is not None:
            self._chk(les_prof, "les_prof", torch.float64, (5, ncol, nk))
        if ps_les is not None:
            self._chk(ps_les, "ps_les", dtype, (ncol,))
        o = _abi.LesForcing()
        res = {}
        alloc = self._alloc_into(res, o, out)
        alloc("ql_ref", (ncol, nk), dtype)
        alloc("ps", (ncol,), dtype)
        if les_prof is not None:
            for n in ("f_u", "f_v", "f_thl", "f_qt", "f_ql"):
                alloc(n, (ncol, nk), dtype)
        if ps_les is not None:
            alloc("f_ps", (ncol,), dtype)
        if want_state:
            for n in ("u", "v", "thl", "qt"):
                alloc(n, (ncol, nk), dtype)
        if couple_surface:
            for n in ("z0m", "z0h", "wthl", "wqt"):
                alloc(n, (ncol,), dtype)
        if diagnostics:
            for n in ("Tv", "THL", "QT", "Zf"):
                alloc(n, (ncol, nlev), dtype)
            alloc("Zh", (ncol, nlev + 1), dtype)
        if want_bracket:
            alloc("bracket", (ncol, nk), torch.int32)
        if zh is not None:
            alloc("slab_idx", (ncol, nlev), torch.int32)
        _abi.check(self._lib.spc_gcm_to_les(self._h, C.byref(s), _ptr(zf), _ptr(zh), nk, _ptr(les_prof), _ptr(ps_les),
                                            float(dt), float(factor), int(bool(couple_surface)), C.byref(o),
                                            self._stream()), "spc_gcm_to_les")
        self.launches += 1
        return res

    # ------------------------------------------------------------------ K3
    def les_to_gcm(self, gcm, zf, zh, slab, aux, slab_idx=None, dt=900.0, factor=1.0, conservative=False,
                   A=None, diagnostics=False, tend_out=None, peer_ptrs=None, peer_col0=0, remote=None, out=None):
        """set_gcm_tendencies for all columns (spcpl.py:388-555) incl. the projected cloud fraction.

        slab: result of slab_reduce (or a dict with 'prof'); aux: dict QL_ice, T (+Rhobf) [ncol,nk].
        remote: RemoteTargets - K3 also stores the block into peer GPUs / pinned host memory and runs the completion
        protocol (peer_ptrs / peer_col0: shorthand for one buffer set without sync).
        Returns dict(tend=[ncol,7,nlev], named views f_T.., A_d, start_index, ...)."""
        s, ncol, nlev, dtype = self._gcm_struct(gcm, False)
        nk = zf.shape[0]
        self._chk(zf, "zf", torch.float64, (nk,))
        if zh is not None:
            self._chk(zh, "zh", torch.float64, (nk,))
        lp = _abi.LesProf()
        lp.prof = self._chk(slab["prof"], "slab.prof", torch.float64, (5, ncol, nk)).data_ptr()
        lp.QL_ice = self._chk(aux["QL_ice"], "QL_ice", dtype, (ncol, nk)).data_ptr()
        lp.T = self._chk(aux["T"], "T", dtype, (ncol, nk)).data_ptr()
        if conservative:
            lp.Rhobf = self._chk(aux["Rhobf"], "Rhobf", dtype, (ncol, nk)).data_ptr()
        ijk_mask = False
        if A is not None:
            lp.A = self._chk(A, "A", dtype, (ncol, nlev)).data_ptr()
        elif slab.get("mask") is not None:
            if slab_idx is None:
                raise ValueError("slab_idx (from gcm_to_les) is needed to project the cloud mask")
            mw = self.mask_words_per_column(slab["dtype"], slab["layout"], slab["nx"], slab["ny"], nk)
            lp.mask = self._chk(slab["mask"], "slab.mask", torch.int32, (ncol, mw)).data_ptr()
            lp.slab_idx = self._chk(slab_idx, "slab_idx", torch.int32, (ncol, nlev)).data_ptr()
            if slab.get("cnt") is not None:
                lp.cnt = self._chk(slab["cnt"], "slab.cnt", torch.int32, (ncol, nk)).data_ptr()
            lp.vol_dtype, lp.layout, lp.nx, lp.ny = _DT[slab["dtype"]], slab["layout"], slab["nx"], slab["ny"]
            ijk_mask = slab["layout"] == _abi.LAYOUT_IJK
        o = _abi.GcmTend()
        if tend_out is None and out and out.get("tend") is not None:
            tend_out = out["tend"]
        tend = tend_out if tend_out is not None else self._empty((ncol, 7, nlev), dtype)
        self._chk(tend, "tend", dtype, (ncol, 7, nlev))
        o.tend = tend.data_ptr()
        o.n_bufs = 1
        if remote is None and peer_ptrs:
            remote = RemoteTargets([list(peer_ptrs)], peer_col0)
        if remote is not None:      # K3 also stores into remote buffers (NVLink peers / pinned host) and signals completion
            remote.fill(o)
        res = {"tend": tend}
        alloc = self._alloc_into(res, o, out)
        alloc("A_d", (ncol, nlev), dtype)
        alloc("start_index", (ncol,), torch.int32)
        if lp.mask:
            alloc("cntslab", (ncol, nlev), torch.int32)
        if diagnostics:
            alloc("t", (ncol, nk), dtype)
            alloc("bracket", (ncol, nlev), torch.int32)
            alloc("bracket_pf", (ncol, nk), torch.int32)
        _abi.check(self._lib.spc_les_to_gcm(self._h, C.byref(s), _ptr(zf), _ptr(zh), nk, C.byref(lp), float(dt),
                                            float(factor), int(bool(conservative)), C.byref(o), self._stream()),
                   "spc_les_to_gcm")
        self.launches += 2 if ijk_mask else 1     # IJK mask: its projection kernel + K3; KJI: the projection is K3's prologue
        for i, n in enumerate(TENDENCIES):
            res[n] = tend[:, i, :]
        return res

    def cloud_fraction(self, slab, slab_idx, dtype=None):
        """les.get_cloudfraction(indices) for all columns (spcpl.py:28,765): (A, cntslab) in ascending
        slab order from the slab_reduce cloud mask."""
        ncol, nlev = slab_idx.shape
        nk = slab["prof"].shape[2]
        dtype = dtype if dtype is not None else slab["dtype"]
        self._chk(slab_idx, "slab_idx", torch.int32)
        if slab.get("mask") is None:
            raise ValueError("slab_reduce was run without want_mask")
        A = self._empty((ncol, nlev), dtype)
        cs = self._empty((ncol, nlev), torch.int32)
        _abi.check(self._lib.spc_cloud_fraction(self._h, _ptr(slab["mask"]), _ptr(slab_idx), _ptr(slab.get("cnt")),
                                                _DT[slab["dtype"]],
                                                slab["layout"], slab["nx"], slab["ny"], nk, ncol, nlev, _DT[dtype],
                                                _ptr(cs), _ptr(A), self._stream()), "spc_cloud_fraction")
        self.launches += 1
        return A, cs

    # ------------------------------------------------------------------ sputils helpers
    def interp(self, x, xp, fp, want_bracket=False):
        """numpy.interp over a batch of rows (sputils.py:82-86). x: [nx] or [nb,nx]; xp, fp: [nb,np]."""
        dtype = xp.dtype
        nb, np_ = xp.shape
        self._chk(xp, "xp", dtype)
        self._chk(fp, "fp", dtype, (nb, np_))
        batched = x.dim() == 2
        nx = x.shape[-1]
        self._chk(x, "x", dtype, (nb, nx) if batched else (nx,))
        out = self._empty((nb, nx), dtype)
        br = self._empty((nb, nx), torch.int32) if want_bracket else None
        _abi.check(self._lib.spc_interp(self._h, _DT[dtype], _ptr(x), int(batched), _ptr(xp), _ptr(fp), nb, nx, np_,
                                        _ptr(out), _ptr(br), self._stream()), "spc_interp")
        self.launches += 1
        return (out, br) if want_bracket else out

    def searchsorted(self, a, v, side="left"):
        """numpy.searchsorted over a batch of rows (sputils.py:88-91). a: [nb,na]; v: [nv] or [nb,nv]."""
        dtype = a.dtype
        nb, na = a.shape
        self._chk(a, "a", dtype)
        batched = v.dim() == 2
        nv = v.shape[-1]
        self._chk(v, "v", dtype, (nb, nv) if batched else (nv,))
        out = self._empty((nb, nv), torch.int32)
        _abi.check(self._lib.spc_searchsorted(self._h, _DT[dtype], _ptr(a), _ptr(v), int(batched), nb, na, nv,
                                              int(side == "right"), _ptr(out), self._stream()), "spc_searchsorted")
        self.launches += 1
        return out

    def exner(self, p, inverse=False):
        """(p/pref0)^(+-rd/cp) (sputils.py:28-34)."""
        self._chk(p, "p")
        out = torch.empty_like(p)
        _abi.check(self._lib.spc_exner(self._h, _DT[p.dtype], _ptr(p), p.numel(), int(inverse), _ptr(out),
                                       self._stream()), "spc_exner")
        self.launches += 1
        return out

    INT_C, INT_RHO, INT_PLAIN, INT_WEIGHTED = 0, 1, 2, 3

    def interp_c(self, Zh, z, q=None, w=None, mode=0):
        """sputils.interp_c / interp_rho / integral over a batch of columns (sputils.py:94-197).
        Zh [nb,nlev+1] descending layer edges; z [nz] float64 ascending cell edges; q, w [nb,nq>=nz-1] cell values.
        mode: INT_C (mass-weighted mean per layer, 0 above z[-1]), INT_RHO (layer-mean of w), INT_PLAIN (integral of q),
        INT_WEIGHTED (integral(q w)/integral(w) without the range rule). Returns [nb,nlev]."""
        nb, nl1 = Zh.shape
        dtype = Zh.dtype
        self._chk(Zh, "Zh", dtype)
        nz = z.shape[0]
        self._chk(z, "z", torch.float64, (nz,))
        ref = q if q is not None else w
        nq = ref.shape[1]
        for t, name in ((q, "q"), (w, "w")):
            if t is not None:
                self._chk(t, name, dtype, (nb, nq))
        out = self._empty((nb, nl1 - 1), dtype)
        _abi.check(self._lib.spc_interp_c(self._h, _DT[dtype], _ptr(Zh), _ptr(z), nz, _ptr(q), _ptr(w), nq, nb, nl1 - 1,
                                          int(mode), _ptr(out), self._stream()), "spc_interp_c")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ set_les_state
    def set_les_state(self, prof, amp, stream_id, nx, ny, seed=42, col0=0, sub=None, clamp0=False,
                      dtype=torch.float32, out=None):
        """Profile -> [ncol,nk,ny,nx] volume with uniform noise (spcpl.py:274-294), Philox4x32-10."""
        ncol, nk = prof.shape
        self._chk(prof, "prof", torch.float64)
        if sub is not None:
            self._chk(sub, "sub", torch.float64, (ncol, nk))
        vol = out if out is not None else self._empty((ncol, nk, ny, nx), dtype)
        self._chk(vol, "vol", dtype, (ncol, nk, ny, nx))
        _abi.check(self._lib.spc_set_les_state(self._h, _ptr(prof), float(amp), int(stream_id), int(seed), int(col0),
                                               _ptr(sub), int(bool(clamp0)), _ptr(vol), _DT[dtype], ncol, nx, ny, nk,
                                               self._stream()), "spc_set_les_state")
        self.launches += 1
        return vol


    # ------------------------------------------------------------------ variability nudge
    def variability_nudge(self, qt, prof, ql_ref, DT, qsat=None, qsat_prof=None, R=None, constant_T=False,
                          thl=None, ql=None, presf=None):
        """spcpl.variability_nudge for all columns (spcpl.py:613-744); qt (and thl) [ncol,nk,ny,nx] are
        updated in place. Returns dict(beta, alpha, qt_std [ncol,nk] f64, status [ncol,nk] i32)."""
        ncol, nk, ny, nx = qt.shape
        dtype = qt.dtype
        self._chk(qt, "qt", dtype)
        self._chk(prof, "prof", torch.float64, (5, ncol, nk))
        self._chk(ql_ref, "ql_ref", dtype, (ncol, nk))
        io = _abi.NudgeIO()
        io.qt, io.prof, io.ql_ref = qt.data_ptr(), prof.data_ptr(), ql_ref.data_ptr()
        if qsat is not None:
            io.qsat = self._chk(qsat, "qsat", dtype, qt.shape).data_ptr()
        if qsat_prof is not None:
            io.qsat_prof = self._chk(qsat_prof, "qsat_prof", dtype, (ncol, nk)).data_ptr()
        if R is not None:
            io.R = self._chk(R, "R", torch.float64, (ncol, ny, nx)).data_ptr()
        if constant_T:
            io.thl = self._chk(thl, "thl", dtype, qt.shape).data_ptr()
            io.ql = self._chk(ql, "ql", dtype, qt.shape).data_ptr()
            io.presf = self._chk(presf, "presf", dtype, (ncol, nk)).data_ptr()
        out = {k: self._empty((ncol, nk), torch.float64) for k in ("beta", "alpha", "qt_std")}
        out["status"] = self._empty((ncol, nk), torch.int32)
        _abi.check(self._lib.spc_variability_nudge(self._h, C.byref(io), _DT[dtype], ncol, nx, ny, nk, float(DT),
                                                   int(bool(constant_T)), _ptr(out["beta"]), _ptr(out["alpha"]),
                                                   _ptr(out["qt_std"]), _ptr(out["status"]), self._stream()),
                   "spc_variability_nudge")
        self.launches += 1
        return out


_default = {}


def default_coupler(device=None):
    """Process-wide Coupler per device."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if dev not in _default:
        _default[dev] = Coupler(dev)
    return _default[dev]
