"""Regenerates tests/golden/realdata_digest.json: order-independent digests of every top-level
lsp*_ng call (problem geometry, score, corners) that the UNMODIFIED reference makes on its own
sample data, harvested by oracle/_ref/spaln_harvest (needs `make -C oracle ref dropin`, i.e.
/root/reference).  tests/test_gpu_realdata.py checks the live harvest on the GPU box against these.

    python tests/golden/make_realdata_digest.py [n_cdna]
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import realdata                     # noqa: E402
import test_gpu_realdata as T       # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    ws = realdata.Workspace()
    runs = {}
    prot = realdata.SEQDB / "dictdisc.faa"
    cq = ws.head_fasta(realdata.SEQDB / "dictdisc.cf", n)
    for tag, opts, query in T.harvest_runs(ws, prot, cq):
        hv = ws.dir / f"{tag}.harvest"
        ws.run("spaln_harvest", opts, query, harvest=hv)
        rows = [(realdata.call_key(c), int(c["score"]), c["skl"].tobytes())
                for kind, c in realdata.read_harvest(hv) if kind == "call"]
        hv.unlink()
        runs[tag] = {"calls": len(rows), "digest": T.digest_of(rows)}
        print(tag, runs[tag])
    (ROOT / "tests" / "golden" / "realdata_digest.json").write_text(
        json.dumps({"n_cdna": n, "runs": runs}, indent=1) + "\n")
    ws.close()


if __name__ == "__main__":
    main()
