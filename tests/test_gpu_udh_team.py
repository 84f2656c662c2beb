"""GPU: the team class of the Hirschberg pass (dp_udh_kernel<..., NW = WARPS_PER_CTA>: a CTA per
problem, 32 strips on one systolic chain) on every problem size.

The device gives a problem a whole CTA from UDH_TEAM_ROWS = 512 query rows on, in batches that cannot fill the device
(spaln_b200/csrc/gspaln_udh.cuh); the seeded and golden Hirschberg tests of test_gpu_parity.py reach
that class only with their longer queries.  Here the same tests run once more in a child process
with GSPALN_UDH_TEAM_ROWS=40 (the library reads it once, when it first plans a batch), so that the
reference fixtures (hirschbergS1_wip, the lspS_ng driver at small -V) and the oracle-checked
seeded problems -- global and local mode, every end-gap flag set, 1 to 8 intermediate rows,
checkpoint re-basing -- all go through the team kernel.  Bar: unchanged, bit-exact."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_hirschberg_and_driver_tests_with_every_problem_in_the_team_class():
    env = dict(os.environ, GSPALN_UDH_TEAM_ROWS="40")
    cmd = [sys.executable, "-m", "pytest", str(ROOT / "tests" / "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
           "-p", "no:cacheprovider", "-k", "hirschberg_wip or lsp_driver or lsp_packed"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=str(ROOT))
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
