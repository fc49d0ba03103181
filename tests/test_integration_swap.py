"""INTEGRATION.md section 1 executed: the real postpic (from /root/reference, which exists only in the build container)
calling postpic_b200's drop-in through the reference's own `_routines.histogramdd` and `MultiSpecies.createField`."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/postpic"), reason="the reference checkout is not on this machine")
def test_reference_routines_run_on_top_of_the_drop_in():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "integration_swap_worker.py")], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "INTEGRATION_SWAP_OK" in out.stdout
