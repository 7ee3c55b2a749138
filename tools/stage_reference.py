"""Stage the UNMODIFIED reference (cdglissov/recurrent-flows-msc) into git-ignored ``baseline/_ref/``.

Run in the build container, where ``/root/reference`` exists (``__graft_entry__.build()`` calls it); the GPU box has
no ``/root/reference``, but ``baseline/_ref/`` travels with the snapshot (git-ignored, not gpurun-ignored), so
``bench.py --impl reference`` and the RFN-level drop-in tests can import the reference's own modules there.

  1. ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of the reference>``:
     its ``setup.py`` uses ``find_packages()``, which picks up ``Flow`` and ``Utils`` (the packages with an
     ``__init__.py``).  The install runs from a /tmp copy because /root/reference is read-only.
  2. ``RFN/RFN_new.py``, ``SRNN/SRNN.py`` (the callers of the hot path, imported as namespace packages) and
     ``main_rfn.py`` (argparse defaults = configuration D) are copied next to them.

Nothing under ``baseline/_ref`` is tracked by git and nothing in the product imports it.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RFMSC_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
EXTRA = ["RFN/RFN_new.py", "RFN/default_rfn_job.sh", "SRNN/SRNN.py", "main_rfn.py", "main_srnn.py"]


def stage(force=False):
    """Returns the staged directory, or None when the reference checkout is absent (GPU box: use what was shipped)."""
    if not os.path.isdir(os.path.join(REF, "Flow")):
        return DST if os.path.isdir(os.path.join(DST, "Flow")) else None
    stamp = os.path.join(DST, ".staged")
    if os.path.exists(stamp) and not force:
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns("Temporary code", "Notebooks", "__pycache__", ".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DST, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:   # fall back to copying the two packages find_packages() would have installed
            sys.stderr.write("stage_reference: pip install failed, copying packages instead:\n" + r.stderr[-2000:] + "\n")
            for pkg in ("Flow", "Utils"):
                shutil.copytree(os.path.join(src, pkg), os.path.join(DST, pkg), dirs_exist_ok=True)
        for rel in EXTRA:
            s = os.path.join(src, rel)
            if os.path.exists(s):
                d = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copy2(s, d)
    shutil.rmtree(os.path.join(DST, "data_generators"), ignore_errors=True)   # needs imageio/torchfile; not on the path
    open(stamp, "w").write("staged from %s\n" % REF)
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
