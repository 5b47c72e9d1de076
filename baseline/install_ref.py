#!/usr/bin/env python
"""Install the UNMODIFIED reference under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

The reference (tylerrleee/mcmc-gpu) is a plain script collection without setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` has nothing to build ("does not appear to be a Python project");
the install is therefore a verbatim copy of its Python sources.  Only runs where /root/reference exists (the build
container); on the GPU box the already-installed copy is used.  Nothing under baseline/_ref is imported by the product:
bench.py's reference arm and cpu_baseline leg run it in a separate process (baseline/run_reference.py).
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("GMC_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["largeScaleChain_multiprocessing.py", "smallScaleChain_multiprocessing.py", "LICENSE"]


def install(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"install_ref: {SRC} not present; keeping {DST} as is ({'present' if os.path.isdir(DST) else 'absent'})")
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(SRC, "gstatsMCMC"), os.path.join(DST, "gstatsMCMC"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb"))
    for f in FILES:
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"install_ref: copied {n} files from {SRC} to {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
