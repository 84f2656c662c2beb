#!/usr/bin/env python
"""Build-container tool (needs oracle/_ref, i.e. /root/reference): shows that the REFERENCE's own
double-affine Hirschberg route corrupts its heap, which is why gspaln_lsp keeps reporting
GSPALN_ST_UNSUPPORTED for it.

    MALLOC_CHECK_=3 python tests/tools/ref_dagp_udh_heapcheck.py udh    # hirschbergS1_wip called directly
    MALLOC_CHECK_=3 python tests/tools/ref_dagp_udh_heapcheck.py lsp    # the whole driver lspS_ng

Options: -Q0 -A2 -S1 -yX0 -yl3 -V256K -TDictyost (PwdB::Noll == 3, small -V).  Observed here (AVX2 build):
the process dies after a handful of planted genes of 60-900 nt -- a segmentation fault, or glibc's
"corrupted size vs. prev_size" / "malloc(): largebin double linked list corrupted" abort; the same script with -A0 instead of -A2 (scalar
hirschbergS_ng, which is on the device) runs clean."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import ref_harness as R  # noqa: E402
from spaln_b200 import workload as synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "lsp"
opts = sys.argv[2] if len(sys.argv) > 2 else "-Q0 -A2 -S1 -yX0 -yl3 -V256K -TDictyost"
ref = R.Reference(opts)
p = ref.params()
rng = np.random.default_rng(19)
for i in range(30):
    g, q, _ = synth.plant_gene(rng, qlen_range=(60, 900), flank=(50, 900))
    t = ref.task(g, q)
    lw, up = t.stripe(p["sh"])
    ex = t.export()
    m = ex["a_right"] - ex["a_left"]
    if which == "lsp":
        r = t.lsp(lw, up)
        print(i, "lspS_ng", r["score"], len(r["skl"]), flush=True)
    elif m >= 16:
        width = up - lw + 3
        mode = 2 if (max(abs(lw), up) + width) < 32767 else 4
        r = t.kernel(lw, up, 2, n_imd=max(1, min(3, m // 16)), mode=mode)
        print(i, "hirschbergS1_wip", r["score"], r["ranges"], flush=True)
    t.close()
print("done: no abort in 30 problems")
