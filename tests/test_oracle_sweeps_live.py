"""CPU, only where `oracle/_ref` is built (the build container): short runs of the sweep tools that
pin the round-2 oracle restatements against the LIVE unmodified reference -- the scalar Hirschberg
passes of the default mode -A0, the -A0 drivers and Cip_score.  One child process per option
string (the reference keeps its options in globals).  The long runs of the same tools are recorded
under profiles/r02_oracle_vs_reference_*.txt."""
import subprocess
import sys
from pathlib import Path

import pytest

import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref not built")
TOOLS = Path(__file__).resolve().parent / "tools"

CASES = [
    ("sweep_oracle_scalar_udh.py", ["dna", "24", "3", "-Q0 -A0 -S1 -yX0 -LS -TDictyost"]),
    ("sweep_oracle_scalar_udh.py", ["prot", "18", "3", "-Q0 -A0 -yX0 -TDictyost"]),
    ("sweep_oracle_lsp.py", ["dna", "30", "3", "-Q0 -A0 -S1 -yX0 -V64K -TDictyost"]),
    ("sweep_oracle_lsp.py", ["prot", "20", "3", "-Q0 -A0 -yX0 -V64K -TDictyost", "cip"]),
    ("sweep_oracle_lsp.py", ["dna", "30", "3", "-Q0 -A2 -S1 -yX0 -V64K -LS -TDictyost", "cip"]),
]


@pytest.mark.parametrize("tool,args", CASES, ids=[f"{c[0][13:-3]}-{'-'.join(c[1][:1] + c[1][3:])[:40]}" for c in CASES])
def test_sweep_against_live_reference(tool, args):
    r = subprocess.run([sys.executable, str(TOOLS / tool)] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-600:], r.stderr[-600:])
    assert ": 0 mismatches in " in r.stdout, r.stdout[-600:]
