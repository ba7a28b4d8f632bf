"""ctypes binding of the C ABI in include/spcpl_b200.h (libspcpl_b200.so, built by build.py).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SPCPL_B200_LIB: alternative build of the same ABI (tools/*_probe.py point it at libspcpl_b200_tune.so, the -DSPC_TUNING
# build that adds the sweep variants and per-handle tuning setters); the package itself always loads the production library
LIB_PATH = os.environ.get("SPCPL_B200_LIB") or os.path.join(HERE, "lib", "libspcpl_b200.so")

SPC_F32, SPC_F64 = 0, 1
LAYOUT_KJI, LAYOUT_IJK = 0, 1
NFIELDS, NTEND = 5, 7
ABI_VERSION = 2
SYNC_EPOCH, SYNC_DONE, SYNC_ERROR, SYNC_FLAG0, SYNC_MAX_SLOTS, SYNC_WORDS = 0, 1, 2, 8, 32, 64
MAX_PEERS = 16

# every symbol include/spcpl_b200.h declares (checked by tests/test_abi.py against the header)
SYMBOLS = ["spc_abi_version", "spc_last_error", "spc_create", "spc_destroy", "spc_mask_words_per_column",
           "spc_slab_reduce", "spc_gcm_to_les", "spc_les_to_gcm", "spc_cloud_fraction", "spc_interp", "spc_searchsorted", "spc_exner", "spc_interp_c",
           "spc_set_les_state", "spc_variability_nudge", "spc_host_register", "spc_host_unregister", "spc_host_device_pointer"]

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double


class GcmCols(C.Structure):
    """struct spc_gcm_cols"""
    _fields_ = [("ncol", _i), ("nlev", _i), ("dtype", _i)] + [(n, _vp) for n in (
        "U", "V", "T", "SH", "QL", "QI", "Pfull", "A", "Zgfull", "Phalf", "Zghalf",
        "Z0M", "Z0H", "QLflux", "QIflux", "SHflux", "TLflux", "TSflux")]


class LesForcing(C.Structure):
    """struct spc_les_forcing"""
    _fields_ = [(n, _vp) for n in (
        "f_u", "f_v", "f_thl", "f_qt", "f_ql", "ql_ref", "u", "v", "thl", "qt", "f_ps", "ps",
        "z0m", "z0h", "wthl", "wqt", "Tv", "THL", "QT", "Zf", "Zh", "bracket", "slab_idx")]


class LesProf(C.Structure):
    """struct spc_les_prof"""
    _fields_ = [(n, _vp) for n in ("prof", "QL_ice", "T", "Rhobf", "A", "mask", "slab_idx", "cnt")] + \
               [(n, _i) for n in ("vol_dtype", "layout", "nx", "ny")]


class GcmTend(C.Structure):
    """struct spc_gcm_tend"""
    _fields_ = [(n, _vp) for n in ("tend", "t", "A_d", "cntslab", "bracket", "bracket_pf", "start_index", "tend_peers")] + \
               [("n_peers", _i), ("peer_col0", _i), ("n_bufs", _i), ("sync", _vp), ("signal", _vp), ("n_signal", _i),
                ("sync_slot", _i), ("n_wait", _i)]


class NudgeIO(C.Structure):
    """struct spc_nudge_io"""
    _fields_ = [(n, _vp) for n in ("qt", "qsat", "qsat_prof", "thl", "ql", "prof", "ql_ref", "presf", "R")]


_lib = None


def lib():
    """Load libspcpl_b200.so once; raise loudly if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "sp_coupler_b200: CUDA library %s is missing. Build it with `python -m sp_coupler_b200.build` "
            "(there is no CPU fallback)." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.spc_abi_version.restype = _i
    L.spc_last_error.restype = C.c_char_p
    L.spc_create.argtypes = [C.POINTER(_vp), _i]
    L.spc_destroy.argtypes = [_vp]
    L.spc_mask_words_per_column.restype = C.c_size_t
    L.spc_mask_words_per_column.argtypes = [_i, _i, _i, _i, _i]
    L.spc_slab_reduce.argtypes = [_vp, C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp]
    L.spc_gcm_to_les.argtypes = [_vp, C.POINTER(GcmCols), _vp, _vp, _i, _vp, _vp, _d, _d, _i,
                                 C.POINTER(LesForcing), _vp]
    L.spc_les_to_gcm.argtypes = [_vp, C.POINTER(GcmCols), _vp, _vp, _i, C.POINTER(LesProf), _d, _d, _i,
                                 C.POINTER(GcmTend), _vp]
    L.spc_cloud_fraction.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]
    L.spc_interp.argtypes = [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]
    L.spc_searchsorted.argtypes = [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]
    L.spc_exner.argtypes = [_vp, _i, _vp, C.c_size_t, _i, _vp, _vp]
    L.spc_interp_c.argtypes = [_vp, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp]
    L.spc_set_les_state.argtypes = [_vp, _vp, _d, C.c_uint32, C.c_uint32, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp]
    L.spc_variability_nudge.argtypes = [_vp, C.POINTER(NudgeIO), _i, _i, _i, _i, _i, _d, _i, _vp, _vp, _vp, _vp, _vp]
    L.spc_host_register.argtypes = [_vp, _vp, C.c_size_t, C.POINTER(_vp)]
    L.spc_host_unregister.argtypes = [_vp, _vp]
    L.spc_host_device_pointer.argtypes = [_vp, _vp, C.POINTER(_vp)]
    if L.spc_abi_version() != ABI_VERSION:
        raise RuntimeError("sp_coupler_b200: %s has ABI version %d, this package needs %d: rebuild it with "
                           "`python -m sp_coupler_b200.build --force`" % (LIB_PATH, L.spc_abi_version(), ABI_VERSION))
    for name in SYMBOLS:
        f = getattr(L, name)
        if f.restype is C.c_int and name not in ("spc_abi_version",):
            f.restype = _i
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().spc_last_error()
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else ""))
