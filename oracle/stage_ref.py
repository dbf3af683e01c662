"""Stage the reference's own Python sources for the CPU arm of bench.py (build container only).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.   Usage:  python -m oracle.stage_ref

The reference (darren-huang/SmartStartContinuous) is pure Python, so "building" it for the CPU
baseline means making its package importable where the benchmark runs: this recipe packs the
``*.py`` files of ``/root/reference/smartstart`` (nothing else: no data, no models) into ONE build
artefact, the git-ignored ``oracle/_ref/smartstart_ref.zip``.  The directory is not gpurun-ignored,
so the archive travels to the GPU box exactly like the built ``libss_b200.so``;
``oracle/ref_harness.py`` unpacks it into a temp dir and imports the reference from there (with stub
modules for TensorFlow / gym / ...) when ``/root/reference`` itself is absent, and ``bench.py --impl reference`` then times the reference's OWN
``NND_MB_agent.get_best_sim_actions`` (``cpu_baseline.kind = "reference"``).  Nothing under
``oracle/_ref`` is ever committed, imported by the product package, or modified.
"""
from __future__ import annotations

import os
import sys
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SS_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
ARCHIVE = os.path.join(DST, "smartstart_ref.zip")


def stage(verbose=True):
    src = os.path.join(SRC, "smartstart")
    if not os.path.isdir(src):
        if verbose:
            print("stage_ref: %s not present, nothing staged" % src)
        return False
    os.makedirs(DST, exist_ok=True)
    n = 0
    tmp = ARCHIVE + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for base, dirs, files in os.walk(src):
            dirs.sort()
            rel = os.path.relpath(base, src)
            for f in sorted(files):
                if f.endswith(".py"):
                    z.write(os.path.join(base, f), os.path.normpath(os.path.join("smartstart", rel, f)))
                    n += 1
    os.replace(tmp, ARCHIVE)
    with open(os.path.join(DST, "STAGED_FROM"), "w") as fh:
        fh.write("%s (%d python files, unmodified; see oracle/stage_ref.py)\n" % (src, n))
    if verbose:
        print("stage_ref: %d files -> %s" % (n, ARCHIVE))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
