"""TEST / BASELINE INFRASTRUCTURE -- installs the UNMODIFIED reference into oracle/_ref.

    python oracle/build_ref.py            (build container only)

The reference (uio-bmi/graph_peak_caller 1.2.3) is pure Python, so "building" it is
a plain ``pip install --no-deps --target oracle/_ref`` of /root/reference, done from a
scratch copy because setuptools writes build/ and *.egg-info into the source tree and
/root/reference is read-only.  ``--no-deps``: its third-party imports
(offsetbasedgraph, pyvg, Bio, pyfaidx) are not in the wheelhouse; oracle/refshim stands
in for them.  Nothing of the reference is copied into tracked files: oracle/_ref/ is
git-ignored (and not gpurun-ignored, so it travels to the GPU box, where
``bench.py --impl reference`` times its ``CallPeaks.run``).
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("GPC_REFERENCE_SOURCE", "/root/reference")


def build(force=False):
    """Returns the install directory, or None when the reference source is absent
    (GPU box: the prebuilt directory travels with the snapshot)."""
    marker = os.path.join(TARGET, "graph_peak_caller", "callpeaks.py")
    if os.path.exists(marker) and not force:
        return TARGET
    if not os.path.isdir(os.path.join(SOURCE, "graph_peak_caller")):
        return TARGET if os.path.exists(marker) else None
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(
            ".git", "tests", "benchmarking", "simulations"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index",
               "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", TARGET, src]
        subprocess.run(cmd, check=True)
    return TARGET


if __name__ == "__main__":
    print(build(force=True))
