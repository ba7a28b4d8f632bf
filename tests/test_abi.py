"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol
include/spcpl_b200.h declares. No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "spcpl_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spc_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from sp_coupler_b200 import build, _abi
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    assert sorted(_abi.SYMBOLS) == syms           # the ctypes binding covers the whole header
    assert _abi.lib().spc_abi_version() == 2


def exported_symbols(path):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return sorted({l.split()[-1] for l in out.splitlines() if l.split() and l.split()[-1].startswith("spc_")})


def test_production_library_exports_exactly_the_header():
    """The converse of the test above: nothing named spc_* is exported that the header does not declare (the tuning
    setters live in libspcpl_b200_tune.so only), the production library carries no sweep variants (size bound) and no
    mutable process-global tuning state."""
    import shutil
    from sp_coupler_b200 import build
    if shutil.which("nm") is None:
        pytest.skip("nm not on PATH")
    path = build.build()
    assert exported_symbols(path) == header_symbols()
    assert os.path.getsize(path) < 3 * 1024 * 1024
    import subprocess
    syms = subprocess.run(["nm", "-C", path], capture_output=True, text=True, check=True).stdout
    assert "g_k1_variant" not in syms and "g_ijk_variant" not in syms and "g_k2_threads" not in syms


def test_sass_is_sm100a_with_tma_bulk():
    """The slab reduction must be a Blackwell TMA kernel: UBLKCP (cp.async.bulk) + mbarrier
    transaction arrives in the SASS, built for sm_100a only."""
    import shutil
    import subprocess
    from sp_coupler_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    path = build.build()
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass and "SYNCS.ARRIVE.TRANS64" in sass
    assert "LDS.128" in sass


def test_production_kernels_do_not_spill():
    """The streaming kernels run one CTA per SM with 12-16 warps: a register spill (stack frame) in any of the production
    instantiations would put local-memory traffic on the HBM-bound path. cuobjdump reads it off the built library."""
    import shutil
    import subprocess
    from sp_coupler_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", build.build()], capture_output=True, text=True).stdout
    usage = {}
    name = None
    for line in out.splitlines():
        line = line.strip()
        if line.startswith("Function "):
            name = line[len("Function "):].rstrip(":")
        elif line.startswith("REG:") and name:
            usage[name] = {k: int(v) for k, v in (kv.split(":") for kv in line.split() if kv.split(":")[1].isdigit())}
    production = {
        "slab_reduce_tma_kernelIfNS_4RingILi16ELi8192ELi1ELb0ELi1ELi1E": 128,     # float32: 16 warps -> at most 128 registers
        "slab_reduce_tma_kernelIdNS_4RingILi12ELi8192ELi1ELb0ELi1ELi1E": 168,     # float64: 12 warps
        "slab_reduce_tma_pair_kernelIfLi12E": 168,
        "slab_reduce_ijk_tma_kernelIfLi5ENS_7IjkRingILi8ELi8192ELi3ELi2E": 255,    # 8 warps x 2 stages
        "slab_reduce_ijk_tma_kernelIdLi5ENS_7IjkRingILi8ELi8192ELi3ELi2E": 255,
    }
    for frag, max_regs in production.items():
        hits = [u for n, u in usage.items() if frag in n]
        assert hits, "kernel %s not found in the library" % frag
        for u in hits:
            assert u["STACK"] == 0 and u["LOCAL"] == 0, (frag, u)
            assert u["REG"] <= max_regs, (frag, u)


def test_struct_layouts_match_header_compiled_as_c(tmp_path):
    """The header must be plain C (what a cgo / Fortran-C / ctypes-gen consumer compiles) and every field of the
    ctypes mirror must sit at the offset gcc gives it."""
    import shutil
    import subprocess
    from sp_coupler_b200 import _abi
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    mirrors = {"spc_gcm_cols": _abi.GcmCols, "spc_les_forcing": _abi.LesForcing, "spc_les_prof": _abi.LesProf,
               "spc_gcm_tend": _abi.GcmTend, "spc_nudge_io": _abi.NudgeIO}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "spcpl_b200.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append('  printf("%s.sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  printf("enum %d %d %d %d %d\\n", SPC_F64, SPC_LAYOUT_IJK, SPC_NFIELDS, SPC_NTEND, SPC_INT_WEIGHTED);',
              '  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = dict(l.rsplit(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
               if not l.startswith("enum"))
    for cname, cls in mirrors.items():
        assert ctypes.sizeof(cls) == int(out[cname + ".sizeof"]), cname
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == int(out["%s.%s" % (cname, fname)]), (cname, fname)
    assert (_abi.SPC_F64, _abi.LAYOUT_IJK, _abi.NFIELDS, _abi.NTEND) == (1, 1, 5, 7)


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sp_coupler_b200.coupler import Coupler
    with pytest.raises(RuntimeError):
        Coupler()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sp_coupler_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_torch_extension_builds_and_registers_ops():
    from sp_coupler_b200 import build, torch_ops
    assert os.path.exists(build.build_torch_ext())
    ops = torch_ops.load()
    assert ops.mask_words_per_column(0, 0, 64, 64, 160) == 64 * 64 * 4 // 4096 * 32 * 160
    for name in ("slab_reduce", "gcm_to_les", "les_to_gcm"):
        assert hasattr(ops, name)


def test_sync_protocol_constants_match_the_header():
    """The host mirror of the completion protocol (K3's flags: _abi.SYNC_*) must use the header's word indices."""
    from sp_coupler_b200 import _abi
    src = open(HEADER).read()
    val = lambda name: int(re.search(name + r"\s*=\s*(\d+)", src).group(1))
    assert (_abi.SYNC_EPOCH, _abi.SYNC_DONE, _abi.SYNC_ERROR, _abi.SYNC_FLAG0, _abi.SYNC_MAX_SLOTS, _abi.SYNC_WORDS) == tuple(
        val(n) for n in ("SPC_SYNC_EPOCH", "SPC_SYNC_DONE", "SPC_SYNC_ERROR", "SPC_SYNC_FLAG0", "SPC_SYNC_MAX_SLOTS", "SPC_SYNC_WORDS"))
    assert _abi.MAX_PEERS == int(re.search(r"#define SPC_MAX_PEERS (\d+)", src).group(1))
    assert _abi.SYNC_FLAG0 + _abi.SYNC_MAX_SLOTS <= _abi.SYNC_WORDS
    assert _abi.ABI_VERSION == int(re.search(r"#define SPC_ABI_VERSION (\d+)", src).group(1))


def test_remote_targets_validation():
    """RemoteTargets (where K3 delivers the tendency block) rejects malformed descriptions on the host, before any
    pointer reaches the kernel."""
    import torch
    from sp_coupler_b200 import _abi
    from sp_coupler_b200.coupler import RemoteTargets
    sync = torch.zeros(_abi.SYNC_WORDS, dtype=torch.int32)
    r = RemoteTargets([[0x1000, 0x2000], [0x3000, 0x4000]], col0=8, sync=sync, signal=[0x5000, 0x6000], slot=1, n_wait=2)
    o = _abi.GcmTend()
    r.fill(o)
    assert (o.n_peers, o.n_bufs, o.peer_col0, o.n_signal, o.sync_slot, o.n_wait) == (2, 2, 8, 2, 1, 2)
    assert o.sync == sync.data_ptr()
    with pytest.raises(ValueError):
        RemoteTargets([[1, 2], [3]])                                   # buffer sets of different length
    with pytest.raises(ValueError):
        RemoteTargets([[1], [2]])                                      # two sets alternate by epoch: need a sync block
    with pytest.raises(ValueError):
        RemoteTargets([list(range(1, _abi.MAX_PEERS + 2))])            # too many targets
    with pytest.raises(ValueError):
        RemoteTargets([[1]], sync=torch.zeros(4, dtype=torch.int32))   # short sync block
    with pytest.raises(ValueError):
        RemoteTargets([[1]], sync=torch.zeros(_abi.SYNC_WORDS, dtype=torch.int64))
    empty = RemoteTargets()
    o2 = _abi.GcmTend()
    empty.fill(o2)
    assert o2.n_peers == 0 and o2.n_bufs == 1 and not o2.sync


def test_flag_polling_returns_and_times_out():
    import threading
    import time
    from sp_coupler_b200.pipeline import _spin_until
    box = [0]
    threading.Timer(0.05, lambda: box.__setitem__(0, 3)).start()
    t0 = time.perf_counter()
    _spin_until(lambda: box[0], 3, 5.0, "test flag")
    assert 0.03 < time.perf_counter() - t0 < 2.0
    with pytest.raises(RuntimeError):
        _spin_until(lambda: 0, 1, 0.2, "never set")


def test_product_package_has_no_numpy_twins():
    """Data generators that restate the reference on the CPU (numpy interp of convert_profiles, the Philox twin of the
    set_les_state kernel) live in tests/synth_les.py; the product package neither contains nor imports them."""
    pkg = os.path.join(ROOT, "sp_coupler_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "synth_les" not in src or f == "synth.py", f          # synth.py only mentions it in its docstring
            assert not re.search(r"^\s*(from|import)\s+(synth_les|cases|conftest)\b", src, flags=re.M), f
            assert "np.interp(" not in src and "numpy.interp(" not in src, f
