"""Put an UNMODIFIED copy of the reference's Python sources for the hot path under ``baseline/_ref`` (git-ignored, NOT
gpurun-ignored: it travels to the GPU box so that ``bench.py`` can time the reference's own functions there).

The reference has no packaging (no setup.py / pyproject.toml), so ``pip install --target baseline/_ref /root/reference``
has nothing to build; this is the equivalent file copy of the packages ``muvo/`` and ``data/`` plus ``constants.py``
(only ``*.py`` / ``*.yml`` / ``*.yaml``; byte-identical, checked).  Run by ``__graft_entry__.build()`` whenever the
checkout is present.  Nothing under baseline/_ref is ever imported by the product; only ``oracle/ref_import.py``
(tests, golden generation, bench baselines) reads it.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MUVO_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP = (".py", ".yml", ".yaml")


def install(verbose: bool = False) -> int:
    if not os.path.isfile(os.path.join(SRC, "muvo", "metrics.py")):
        if verbose:
            print(f"install_reference: no checkout at {SRC}; keeping whatever is in {DST}")
        return 0
    n = 0
    for top in ("muvo", "data", "constants.py"):
        s = os.path.join(SRC, top)
        if os.path.isfile(s):
            files = [(s, os.path.join(DST, top))]
        else:
            files = []
            for d, _, names in os.walk(s):
                for nm in names:
                    if nm.endswith(KEEP):
                        p = os.path.join(d, nm)
                        files.append((p, os.path.join(DST, os.path.relpath(p, SRC))))
        for a, b in files:
            os.makedirs(os.path.dirname(b), exist_ok=True)
            if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
                shutil.copyfile(a, b)
            assert filecmp.cmp(a, b, shallow=False)
            n += 1
    with open(os.path.join(DST, "INSTALLED_FROM"), "w") as f:
        f.write(f"{SRC}\nfiles: {n} (unmodified copies; see tools/install_reference.py)\n")
    if verbose:
        print(f"install_reference: {n} files -> {DST}")
    return n


if __name__ == "__main__":
    install(verbose=True)
    sys.exit(0)
