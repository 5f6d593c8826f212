"""Generate oracle trace fixtures at sizes the oracle needs minutes for (run here, committed).

    python tests/golden/make_trace_fixtures.py poisson7:64 poisson27:32 ...

Each case writes tests/golden/trace_<name>_<n>.json.gz: the oracle's per-stage trace (tag, FNV-1a
hash, byte count of every traced intermediate array, amg_oracle.c), the per-level sizes and
smoother parameters, and the hierarchy fingerprint (oracle.hierarchy_hash).  The GPU tests
(tests/test_gpu_parity.py::test_trace_fixture_parity) replay the same input through the CUDA path
and compare all of it.  The oracle is pinned to the unmodified reference on the inputs the
reference can process (tests/test_oracle.py); on finite-difference Poisson/diffusion inputs it
uses the CHECKED sp_add.
"""
import gzip
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from util import orc  # noqa: E402
from omp_amg_b200 import matrices as M  # noqa: E402


def make(name, n):
    O = orc.Oracle()
    mat = M.by_name(name, n)
    t = time.time()
    h = O.setup_raw(*mat, orc.SEQ, trace=True)
    secs = time.time() - t
    tr = O.trace()
    H = O.fetch(h)
    O.free(h)
    out = {"workload": name, "n": n, "mode": "seq", "oracle_seconds": round(secs, 1),
           "nlevels": H.nlevels, "nullspace": H.nullspace,
           "levels": [{"info": [int(x) for x in lev["info"]], "m": lev["m"], "rho": lev["rho"].hex() if hasattr(lev["rho"], "hex") else float(lev["rho"]).hex()}
                      for lev in H.levels],
           "hierarchy_hash": "%016x" % orc.hierarchy_hash(H),
           "trace": [[tag, "%016x" % hs, nb] for tag, hs, nb in tr]}
    path = os.path.join(HERE, "trace_%s_%d.json.gz" % (name, n))
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("%s:%d  %.1f s  levels %s  %d trace records  hash %s -> %s" % (
        name, n, secs, [lev["info"][0] for lev in H.levels], len(tr), out["hierarchy_hash"], path), flush=True)


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        name, n = spec.split(":")
        make(name, int(n))
