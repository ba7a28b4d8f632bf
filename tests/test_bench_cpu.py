"""bench.py legs that run without a GPU: the reference arm prints the contract's JSON line, and the B200 arm
refuses to run (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_prints_the_contract_line():
    """The reference arm times the UNMODIFIED reference (oracle/_ref, made by oracle/make_ref.py where /root/reference
    exists; on the GPU box the copy travels with the snapshot) and keeps the numpy port beside it."""
    p = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-cols", "4",
                        "--ref-procs", "2"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "coupled columns/s" and line["unit"] == "columns/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["scaling"] == "strong" and line["config"]["ncol_total"] == 2048
    from oracle import ref_driver
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_driver.available() else "port")
    assert line["cpu_baseline"]["cores"] == 2
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["cpu_baseline_port"]["kind"] == "port" and line["cpu_baseline_port"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["gpu_launches"] == 0


def test_reference_copy_is_verbatim():
    """oracle/_ref (git-ignored) holds byte-identical copies of the reference modules; the recipe is idempotent."""
    from oracle import make_ref, ref_driver
    m = make_ref.make()
    if m is None:
        if not make_ref.verify():
            pytest.skip("no reference tree and no oracle/_ref on this machine")
        return
    assert make_ref.verify()
    for name, h in m["files"].items():
        assert h == make_ref.sha256(os.path.join(make_ref.SRC, "splib", name)), name
    assert ref_driver.available()
    gi = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in gi


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = subprocess.run([sys.executable, BENCH, "--steps", "1", "--no-cpu"], capture_output=True, text=True, timeout=600)
    assert p.returncode != 0
    assert p.stdout.strip() == ""          # no JSON line from a machine that cannot run the kernels
