"""Regenerates the stored METIS orderings of the large BAL workloads (gtsam_b200/data_bal_*_metis.npz) with the
UNMODIFIED reference's own Ordering::Metis (gtsam/inference/Ordering.cpp:210-255 -> METIS_NodeND), through
oracle/_ref/ref_harness order.  Orderings are INPUTS at the C-ABI boundary; these files are how
bench.py gets the reference's ordering on the GPU box, where the reference tree does not exist.  Run from the repo root
in the build container (about 2 minutes, 6 GB of host memory for the 10M-factor graph):

    python tests/golden/make_orderings.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gtsam_b200 import datasets  # noqa: E402
from oracle import refio  # noqa: E402


def main():
    assert refio.have_ref(), "build oracle/_ref first: make -C oracle ref"
    for name in ("bal_1m_metis", "bal_c4_metis", "bal_c5_metis"):
        fn, kw = datasets.WORKLOADS[name]
        kw = dict(kw)
        kind = kw.pop("ordering")
        prob = fn(**kw)     # Schur-ordered twin: the ordering does not change the graph
        t = time.time()
        o = refio.run("order", prob, kind)["ordering"]
        path = datasets.ordering_file(kw["ncams"], kw["npoints"], kw.get("obs_per_point", 6), kw.get("visibility", "scattered"),
                                      kw.get("seed", 42), kind)
        np.savez_compressed(path, ordering=o.astype(np.int32))
        print(name, len(o), "variables,", "%.1f s," % (time.time() - t), os.path.getsize(path), "bytes ->", path)


if __name__ == "__main__":
    main()
