"""The drop-in property, whole program: the reference's own main.py on its demo, unmodified vs with the INTEGRATION.md shims
(model, training_utils, and the data_utils functions of row f-4) — console output and training log must be identical
(tests/soak/main_py_soak_vs_reference.py; the engine's host program runs over the CPU emulation of the kernels).
Needs the reference checkout, which exists in the build container only: skipped elsewhere."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference checkout exists in the build container only")
def test_main_py_with_the_shims_prints_what_the_reference_prints():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "soak", "main_py_soak_vs_reference.py"), "demo"],
                       capture_output=True, text=True, timeout=1500, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "109 console lines, 0 differing; 78 log-file lines, 0 differing" in r.stdout
    assert "with the data_utils shim as well: 0 console lines and 0 log-file lines differing" in r.stdout
