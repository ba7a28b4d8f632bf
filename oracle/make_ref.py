"""Recipe that makes the UNMODIFIED reference importable where /root/reference does not exist (the GPU box) —
TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            # copies, verifies, prints the manifest

Copies the pure-Python modules of the reference that the coupling path imports
(/root/reference/splib/{__init__,spcpl,sputils,spio,spdummy,haversine}.py) byte for byte into the git-ignored
directory oracle/_ref/splib/ and writes oracle/_ref/MANIFEST.json with their sha256. oracle/_ref/ is listed in
.gitignore (reference sources never enter this repository's history) but not in .gpurunignore, so it travels to the GPU
box with the snapshot exactly like the built .so files do. There, `bench.py --impl reference` and the `cpu_baseline`
leg import it through oracle/ref_driver.py under the unit shim oracle/stubs (the reference needs amuse / omuse /
netCDF4 / shapely, which are not installed) and time the reference's own functions on the box's host cores.
Nothing under sp_coupler_b200/ imports it. __graft_entry__.build() runs this recipe when /root/reference is present.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SPC_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["__init__.py", "spcpl.py", "sputils.py", "spio.py", "spdummy.py", "haversine.py"]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def make(verbose=False):
    """Copy (or refresh) the files; returns the manifest dict, or None when the reference tree is absent."""
    src_dir = os.path.join(SRC, "splib")
    if not os.path.isfile(os.path.join(src_dir, "spcpl.py")):
        return None
    dst_dir = os.path.join(DST, "splib")
    os.makedirs(dst_dir, exist_ok=True)
    manifest = {"source": src_dir, "files": {}}
    for name in FILES:
        s, d = os.path.join(src_dir, name), os.path.join(dst_dir, name)
        if not os.path.exists(d) or sha256(s) != sha256(d):
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
        manifest["files"][name] = sha256(d)
        assert manifest["files"][name] == sha256(s), name
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print(json.dumps(manifest, indent=1, sort_keys=True))
    return manifest


def verify():
    """True when oracle/_ref holds exactly the files its manifest lists, unmodified since they were copied."""
    mpath = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(mpath):
        return False
    m = json.load(open(mpath))
    return all(os.path.exists(os.path.join(DST, "splib", n)) and sha256(os.path.join(DST, "splib", n)) == h
               for n, h in m["files"].items())


if __name__ == "__main__":
    m = make(verbose=True)
    if m is None:
        print("reference tree not found at %s; oracle/_ref %s" % (SRC, "is intact" if verify() else "is absent"))
        sys.exit(0 if verify() else 1)
