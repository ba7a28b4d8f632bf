"""GPU-resident stand-in models with the duck-typed protocol the coupler drives.

Same role and method names as the reference's analytic stand-ins (splib/spdummy.py: dummy_gcm,
dummy_les) plus the methods current spcpl/splib call that the reference dummies lack (get_rain,
get_rhof, get_rhobf, set_tendency_QL, set_*_surf, return_request=...).

* `gpu_les_batch` owns the 3-D LES state of all local columns in HBM ([ncol][nk][ny][nx] per field);
  `gpu_les` objects are per-column views on it (what `les_models` holds in splib).
* `gpu_gcm` is a host-side synthetic GCM (profiles in pinned host memory), as OpenIFS is a CPU code.

The LES "dynamics" is a relaxation stand-in (the applied forcing is integrated over the step and
the noise is kept), enough to close the coupling loop; real LES time stepping is outside the path.
"""
import datetime

import numpy as np
import torch

from . import synth
from .constants import LES_FIELDS
from .coupler import default_coupler
from .pipeline import CouplingPipeline


class Request(object):
    """Completed stand-in of an AMUSE async request (amuse.rfi.async_request)."""

    def __init__(self, value=None):
        self._value = value

    def result(self):
        return self._value

    def is_result_available(self):
        return True

    def wait(self):
        return None


def _ret(value, return_request):
    return Request(value) if return_request else value


class dummy_base(object):
    """spdummy.py:13-56."""

    def __init__(self, nprocs=1):
        self.step = 0
        self.timestep = 900.0                       # s (oifs-input/fort.4:52)
        self.starttime = datetime.datetime(2000, 1, 1)
        self.model_time = 0.0
        self.number_of_workers = nprocs
        self.support_async = True

    def get_timestep(self):
        return self.timestep

    def get_model_time(self):
        return self.model_time

    def get_start_datetime(self):
        return self.starttime

    def cleanup_code(self):
        return True

    def stop(self):
        return True


class gpu_gcm(dummy_base):
    """Synthetic GCM on a lon x lat grid (default 64 x 32 = the 2048 T21 Gaussian points)."""

    def __init__(self, nprocs=1, num_lons=64, num_lats=32, nlev=91, dtype=np.float32, seed=42):
        super(gpu_gcm, self).__init__(nprocs)
        self.num_lons, self.num_lats, self.ktot = num_lons, num_lats, nlev
        self.dtype = dtype
        n = num_lons * num_lats
        lats = 180.0 * (np.arange(num_lats) + 0.5) / num_lats - 90.0
        lons = 360.0 * np.arange(num_lons) / num_lons
        self.latitudes = np.repeat(lats, num_lons)          # spdummy.py:97-100
        self.longitudes = np.tile(lons, num_lats)
        self.state = synth.make_gcm_columns(n, nlev, seed=seed, dtype=dtype)
        self.tendencies = {}
        self.mask = set()
        self.first_half_step_done = False
        self.vdf_in_sp_mask = True

    def get_itot(self):
        return self.num_lons

    def get_jtot(self):
        return self.num_lats

    def get_ktot(self):
        return self.ktot

    def initialize_code(self):
        pass

    def commit_parameters(self):
        pass

    def commit_grid(self):
        pass

    def set_mask(self, i):
        self.mask.add(int(i))

    def set_vdf_in_sp_mask(self, value):
        self.vdf_in_sp_mask = value

    def get_profile_fields(self, name, index):
        """[len(index), nlev(+1)] host array (spdummy.py:119-125)."""
        return self.state[name][np.asarray(index, dtype=np.int64)]

    def get_surface_field(self, name, index):
        return self.state[name][np.asarray(index, dtype=np.int64)]

    def set_profile_tendency(self, field, index, vals):
        """spdummy.py:174-175; values may be device tensors."""
        v = vals.detach().cpu().numpy() if isinstance(vals, torch.Tensor) else np.asarray(vals)
        self.tendencies.setdefault(field, {})[int(index)] = v

    def set_profile_tendencies(self, index, packed, lev0=0):
        """Batched form: packed [ncol][7][nlev - lev0] host tensor in spc tendency order; with a level window
        (lev0 > 0) the block holds the lowest levels only and the tendencies above are zero (spcpl.py:527-533)."""
        from .constants import TENDENCIES
        arr = packed.numpy() if isinstance(packed, torch.Tensor) else np.asarray(packed)
        idx = np.asarray(index, dtype=np.int64)
        for n, name in enumerate(TENDENCIES):
            full = np.zeros((arr.shape[0], self.ktot), dtype=arr.dtype)
            full[:, lev0:] = arr[:, n, :]
            self.tendencies.setdefault(name[2:], {})["batch"] = (idx, full)

    def evolve_model_until_cloud_scheme(self):
        return True

    def evolve_model_cloud_scheme(self):
        self.tendencies = {}                               # "overwrites set tendencies" (splib.py:299)
        return True

    def evolve_model_from_cloud_scheme(self):
        """Apply the LES tendencies to the GCM state over one time step, then advance the clock."""
        for field, rec in self.tendencies.items():
            if field not in self.state:
                continue
            for key, v in rec.items():
                if key == "batch":
                    idx, arr = v
                    self.state[field][idx] += (arr * self.timestep).astype(self.dtype)
                else:
                    self.state[field][key] += (np.asarray(v) * self.timestep).astype(self.dtype)
        self.model_time += self.timestep
        return True

    def evolve_model_single_step(self):
        return self.evolve_model_from_cloud_scheme()


class gpu_les_batch(object):
    """All local LES models: five [ncol][nk][ny][nx] volumes + LES-internal profiles in HBM."""

    def __init__(self, ncol, nlev, nx=64, ny=64, nk=160, dz=25.0, dtype=torch.float32, device=None, seed=42,
                 col0=0, couple_surface=True, group=None, gather=True, ncol_total=None):
        self.cpl = default_coupler(device)
        self.ncol, self.nlev, self.nx, self.ny, self.nk, self.dtype = ncol, nlev, nx, ny, nk, dtype
        self.seed, self.col0 = seed, col0
        self.zf_host, self.zh_host = synth.les_grid(nk, dz)
        self.gather_mode = gather                  # "host": no device gather, tendencies leave through pipeline.HostExchange
        self.pipe = CouplingPipeline(self.cpl, self.zf_host, self.zh_host, ncol, nlev, dtype, couple_surface,
                                     group=group, gather=False if gather == "host" else gather)
        dev = self.cpl.device
        self.vols = [torch.zeros((ncol, nk, ny, nx), dtype=dtype, device=dev) for _ in LES_FIELDS]
        ndt = np.float32 if dtype == torch.float32 else np.float64
        aux = synth.make_les_aux(ncol, nk, seed=seed, dtype=ndt, col0=col0, ncol_total=ncol_total)
        self.aux = {k: torch.from_numpy(v).to(dev) for k, v in aux.items()}
        self.pipe.attach_les(self.vols, self.aux)
        z = lambda *s: torch.zeros(s, dtype=dtype, device=dev)
        self.tend = {f: z(ncol, nk) for f in ("U", "V", "THL", "QT", "QL")}
        self.tend_ps = z(ncol)
        self.ql_ref = z(ncol, nk)
        self.surf = {n: z(ncol) for n in ("z0m", "z0h", "wt", "wq")}
        self.model_time = 0.0
        self.state_version = 0      # bumped whenever the 3-D state changes; keys the per-column slab cache
        self.models = [gpu_les(self, i) for i in range(ncol)]

    def initialize_state(self, profiles):
        """set_les_state for every column: profiles dict u,v,thl,qt [ncol][nk] (device)."""
        amp = synth.NOISE_AMP
        for f, key in (("THL", "thl"), ("QT", "qt"), ("U", "u"), ("V", "v")):
            self.cpl.set_les_state(profiles[key].double().contiguous(), amp[f], synth.STREAM[f], self.nx, self.ny,
                                   seed=self.seed, col0=self.col0, dtype=self.dtype, out=self.vols[LES_FIELDS.index(f)])
        self._diagnose_ql(profiles["qt"].double())
        self.state_version += 1

    def _diagnose_ql(self, qt_prof):
        """Stand-in saturation adjustment: ql = max(qt - qsat, 0) with qsat = <qt> + amp*s(k)."""
        qsat = qt_prof + synth.NOISE_AMP["QT"] * torch.from_numpy(synth.cloud_offset(self.nk)).to(qt_prof.device)[None, :]
        self.qsat = qsat
        qt = self.vols[LES_FIELDS.index("QT")]
        torch.clamp(qt - qsat.to(self.dtype)[:, :, None, None], min=0, out=self.vols[LES_FIELDS.index("QL")])

    def evolve(self, t_end):
        """Relaxation stand-in for LES time stepping: integrate the applied forcings to t_end."""
        dt = float(t_end) - self.model_time
        if dt <= 0:
            return
        for f in ("THL", "QT", "U", "V"):
            self.vols[LES_FIELDS.index(f)].add_((self.tend[f] * dt)[:, :, None, None])
        self.aux["PS"].add_(self.tend_ps * dt)
        qt = self.vols[LES_FIELDS.index("QT")]
        torch.clamp(qt - self.qsat.to(self.dtype)[:, :, None, None], min=0, out=self.vols[LES_FIELDS.index("QL")])
        self.model_time = float(t_end)
        self.state_version += 1


class gpu_les(dummy_base):
    """Per-column view with the LES protocol of spcpl."""

    def __init__(self, batch, i):
        super(gpu_les, self).__init__(1)
        self.batch, self.i = batch, i
        self.grid_index = i
        self.rain = 0.0

    # geometry ---------------------------------------------------------------------------------
    def get_itot(self):
        return self.batch.nx

    def get_jtot(self):
        return self.batch.ny

    def get_ktot(self):
        return self.batch.nk

    def get_zf(self, return_request=False):
        return _ret(self.batch.pipe.zf, return_request)

    def get_zh(self, return_request=False):
        return _ret(self.batch.pipe.zh, return_request)

    def get_model_time(self):
        return self.batch.model_time

    # state ------------------------------------------------------------------------------------
    def volumes(self):
        """The five [1][nk][ny][nx] views of this column."""
        return [v[self.i:self.i + 1] for v in self.batch.vols]

    def _slab(self, want_mask=False):
        """Slab means (+ counts and cloud mask) of this column. The five get_profile_* calls of one step
        (spcpl.py:303-307) and get_cloudfraction share ONE reduction: the result is cached until the 3-D state
        changes (batch.state_version), instead of re-reading all five volumes per call."""
        key = (self.batch.state_version, bool(want_mask))
        c = getattr(self, "_slab_cache", None)
        if c is not None and c[0][0] == key[0] and (c[0][1] or not want_mask):
            return c[1]
        res = self.batch.cpl.slab_reduce(self.volumes(), want_cnt=want_mask, want_mask=want_mask)
        self._slab_cache = (key, res)
        return res

    def get_profile(self, name, return_request=False):
        f = LES_FIELDS.index(name)
        return _ret(self._slab()["prof"][f, 0], return_request)

    def get_profile_U(self, return_request=False):
        return self.get_profile("U", return_request)

    def get_profile_V(self, return_request=False):
        return self.get_profile("V", return_request)

    def get_profile_THL(self, return_request=False):
        return self.get_profile("THL", return_request)

    def get_profile_QT(self, return_request=False):
        return self.get_profile("QT", return_request)

    def get_profile_QL(self, return_request=False):
        return self.get_profile("QL", return_request)

    def _aux(self, name, return_request):
        return _ret(self.batch.aux[name][self.i], return_request)

    def get_profile_QL_ice(self, return_request=False):
        return self._aux("QL_ice", return_request)

    def get_profile_QR(self, return_request=False):
        return self._aux("QR", return_request)

    def get_profile_T(self, return_request=False):
        return self._aux("T", return_request)

    def get_presf(self, return_request=False):
        return self._aux("presf", return_request)

    def get_rhof(self, return_request=False):
        return self._aux("Rhof", return_request)

    def get_rhobf(self, return_request=False):
        return self._aux("Rhobf", return_request)

    def get_surface_pressure(self, return_request=False):
        return self._aux("PS", return_request)

    def get_rain(self, return_request=False):
        return self._aux("Rain", return_request)

    def get_field(self, name):
        """3-D field in the OMUSE (itot, jtot, ktot) view (spcpl.py:627-628)."""
        return self.batch.vols[LES_FIELDS.index(name)][self.i].permute(2, 1, 0)

    def get_cloudfraction(self, indices, return_request=False):
        """Fraction of horizontal points with ql > 0 anywhere in each slab of LES levels delimited
        by `indices` (spcpl.py:28,765), ascending slab order."""
        idx = torch.as_tensor(indices, dtype=torch.int32, device=self.batch.cpl.device).reshape(1, -1).contiguous()
        slab = self._slab(want_mask=True)
        A, _ = self.batch.cpl.cloud_fraction(slab, idx)
        return _ret(A[0], return_request)

    def set_field(self, fid, values):
        """values: (itot, jtot, ktot) array as in spcpl.set_les_state (spcpl.py:288-291)."""
        v = torch.as_tensor(values, dtype=self.batch.dtype, device=self.batch.cpl.device)
        self.batch.vols[LES_FIELDS.index(fid)][self.i].copy_(v.permute(2, 1, 0))
        self.batch.state_version += 1

    def set_surface_pressure(self, value):
        self.batch.aux["PS"][self.i] = float(value)

    # forcings ---------------------------------------------------------------------------------
    def _set(self, store, values, return_request):
        store[self.i].copy_(torch.as_tensor(values, device=store.device).to(store.dtype))
        return _ret(None, return_request)

    def set_tendency_U(self, values, return_request=False):
        return self._set(self.batch.tend["U"], values, return_request)

    def set_tendency_V(self, values, return_request=False):
        return self._set(self.batch.tend["V"], values, return_request)

    def set_tendency_THL(self, values, return_request=False):
        return self._set(self.batch.tend["THL"], values, return_request)

    def set_tendency_QT(self, values, return_request=False):
        return self._set(self.batch.tend["QT"], values, return_request)

    def set_tendency_QL(self, values, return_request=False):
        return self._set(self.batch.tend["QL"], values, return_request)

    def set_tendency_surface_pressure(self, values, return_request=False):
        return self._set(self.batch.tend_ps, values, return_request)

    def set_ref_profile_QL(self, values, return_request=False):
        return self._set(self.batch.ql_ref, values, return_request)

    def set_z0m_surf(self, value, return_request=False):
        return self._set(self.batch.surf["z0m"], value, return_request)

    def set_z0h_surf(self, value, return_request=False):
        return self._set(self.batch.surf["z0h"], value, return_request)

    def set_wt_surf(self, value, return_request=False):
        return self._set(self.batch.surf["wt"], value, return_request)

    def set_wq_surf(self, value, return_request=False):
        return self._set(self.batch.surf["wq"], value, return_request)

    def evolve_model(self, stop_time, exactEnd=True):
        # the batch advances all columns together (first caller does the work)
        self.batch.evolve(stop_time)

    def write_restart(self):
        return True
