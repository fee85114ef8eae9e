"""Regenerates tests/golden/*.bin by running the UNMODIFIED reference
(oracle/_ref/ref_harness, built by oracle/Makefile from /root/reference) on
small seeded problems.  Run from the repo root in the build container:

    python tests/golden/make_golden.py

The fixtures pin the C oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_parity.py).  Each case stores the problem (*.prob.bin, the
gtsam_b200.problem.Problem.save format) and the reference's outputs
(*.dump0.bin: lambda=0; *.dump1.bin: lambda=1e-2 with diagonal damping;
*.lm.bin: LevenbergMarquardtOptimizer trace; *.gn.bin: GaussNewton trace; *.dl.bin: DoglegOptimizer trace; *.marg.bin: Marginals covariances).
"""
import os
import shutil
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gtsam_b200 import datasets  # noqa: E402
from oracle import refio  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
H = refio.HARNESS
REF_DATA = "/root/reference/examples/Data"


def emit(name, prob, lm_iters=30, gn_iters=0, ceres=False, dl_iters=0, marg=False):
    ppath = os.path.join(HERE, f"{name}.prob.bin")
    prob.save(ppath)
    subprocess.check_call([H, "dump", ppath, os.path.join(HERE, f"{name}.dump0.bin"), "0", "0"])
    subprocess.check_call([H, "dump", ppath, os.path.join(HERE, f"{name}.dump1.bin"), "1e-2", "1"])
    subprocess.check_call([H, "lm", ppath, os.path.join(HERE, f"{name}.lm.bin"), str(lm_iters), str(int(ceres))])
    if gn_iters:
        subprocess.check_call([H, "gn", ppath, os.path.join(HERE, f"{name}.gn.bin"), str(gn_iters)])
    if dl_iters:   # DoglegOptimizer trace, deltaInitial = 1
        subprocess.check_call([H, "dogleg", ppath, os.path.join(HERE, f"{name}.dl.bin"), str(dl_iters), "1.0"])
    if marg:       # Marginals::marginalCovariance of every variable
        subprocess.check_call([H, "marginals", ppath, os.path.join(HERE, f"{name}.marg.bin")])
    print("wrote", name, prob.nvars, "vars", prob.nfactors, "factors")


def with_ordering(prob, kind):
    o = refio.run("order", prob, kind)["ordering"]
    prob.ordering = o.copy()
    return prob


def main():
    assert refio.have_ref(), "build oracle/_ref first: make -C oracle ref"
    subprocess.check_call([H, "kat", os.path.join(HERE, "geometry_kat.bin")])
    emit("bal_tiny_s2", datasets.make("bal_tiny"), dl_iters=5, marg=True)
    emit("bal_tiny_body_sensor", datasets.make("bal_tiny", seed=13, body_sensor=True), dl_iters=5)
    emit("bal_tiny_bundler", datasets.make("bal_tiny", camera_model="bundler"), ceres=True, dl_iters=5, marg=True)
    emit("bal_tiny_colamd", with_ordering(datasets.make("bal_tiny", seed=5), "colamd"), dl_iters=5)
    emit("sphere_tiny", datasets.make("sphere_tiny"), gn_iters=3, dl_iters=8, marg=True)
    emit("sphere_small_colamd", with_ordering(datasets.sphere(layers=8, per_ring=12, seed=3), "colamd"), gn_iters=3, dl_iters=8)
    emit("sphere_tiny_gaussian", datasets.sphere(layers=5, per_ring=8, seed=11, noise="gaussian"), gn_iters=2, dl_iters=6, marg=True)
    from gtsam_b200 import problem as Pq
    import numpy as np
    emit("sphere_tiny_huber", datasets.sphere(layers=5, per_ring=8, seed=12, robust=(Pq.ROBUST_HUBER, 1.345)), gn_iters=2, dl_iters=6)
    emit("sphere_tiny_cauchy", datasets.sphere(layers=5, per_ring=8, seed=14, noise="gaussian", robust=(Pq.ROBUST_CAUCHY, 2.0)))
    bt = datasets.make("bal_tiny", seed=15)
    bt.groups[0].robust_kind, bt.groups[0].robust_param = Pq.ROBUST_TUKEY, 4.685
    emit("bal_tiny_tukey", bt)
    bf = datasets.make("bal_tiny", seed=16)
    bf.groups[0].robust_kind, bf.groups[0].robust_param = Pq.ROBUST_FAIR, 1.3998
    emit("bal_tiny_fair", bf)
    # factors of one kind that are NOT consecutive in the graph (explicit graph positions)
    si = datasets.sphere(layers=5, per_ring=8, seed=17)
    nb = si.groups[0].count
    kpos = nb // 3
    si.groups[0].graph_index = np.concatenate([np.arange(kpos), np.arange(kpos + 1, nb + 1)])
    si.groups[1].graph_index = np.array([kpos])
    emit("sphere_tiny_interleaved", si, gn_iters=2)
    emit("sphere_small_metis", with_ordering(datasets.sphere(layers=8, per_ring=12, seed=4), "metis"))
    # BAL eliminated in the reference's METIS nested-dissection order (BASELINE configs[3] names METIS): camera and point
    # cliques interleave, every point is still a leaf
    emit("bal_small_metis", with_ordering(datasets.bal(ncams=12, npoints=300, seed=21), "metis"))
    # BASELINE configs[0]'s factor family: planar pose graphs (BetweenFactor<Pose2> + PriorFactor<Pose2>)
    emit("pose2_ring", datasets.pose2_ring(n=40), gn_iters=4, dl_iters=6, marg=True)
    emit("pose2_ring_colamd", with_ordering(datasets.pose2_ring(n=60, seed=5), "colamd"), gn_iters=4)
    # ---- input formats: synthetic text files written by gtsam_b200.io, parsed by the
    # REFERENCE's loaders (readG2o / SfmData::FromBalFile) into *.prob.bin; tests compare our readers with them
    import numpy as np
    from gtsam_b200 import io, problem as Pm
    ddir = os.path.join(HERE, "data")
    os.makedirs(ddir, exist_ok=True)
    sp = datasets.sphere(layers=4, per_ring=6, seed=21, noise="gaussian")
    R = sp.groups[0].noise.reshape(-1, 6, 6)
    info = np.einsum("nki,nkj->nij", R, R)
    g2o_txt = os.path.join(ddir, "synthetic_sphere.g2o")
    io.write_g2o_3d(g2o_txt, sp.values.reshape(-1, 12), sp.groups[0].keys, sp.groups[0].meas, info)
    subprocess.check_call([H, "g2ofile", g2o_txt, os.path.join(HERE, "synthetic_sphere_g2o.prob.bin")])
    bl = datasets.make("bal_tiny", camera_model="bundler", seed=9)
    nc = int((bl.var_type == Pm.VAR_CAM_BUNDLER).sum())
    bal_txt = os.path.join(ddir, "synthetic_bal.txt")
    io.write_bal(bal_txt, bl.values[:nc * 17].reshape(nc, 17), bl.values[nc * 17:].reshape(-1, 3),
                 bl.groups[0].keys[:, 0], bl.groups[0].keys[:, 1] - nc, bl.groups[0].meas)
    subprocess.check_call([H, "balfile", bal_txt, os.path.join(HERE, "synthetic_bal.prob.bin"), "1"])
    # the reference's own example data through its own parser: Pose3SLAMExample_g2o on pose3example.txt
    p3 = os.path.join(HERE, "pose3example.prob.bin")
    subprocess.check_call([H, "g2ofile", os.path.join(REF_DATA, "pose3example.txt"), p3])
    emit("pose3example", Pm.Problem.load(p3), lm_iters=10, gn_iters=5)
    # BASELINE.json configs[0] (CPU plumbing): the reference's Pose2 g2o example path
    with open(os.path.join(HERE, "config1_pose2slam_g2o.json"), "w") as f:
        f.write(subprocess.check_output([H, "pose2", os.path.join(REF_DATA, "noisyToyGraph.txt")]).decode())
    # GncOptimizer<GncParams<LevenbergMarquardtParams>> on graphs with injected outliers (TLS and GM)
    rng = np.random.default_rng(5)
    so = datasets.sphere(layers=5, per_ring=8, seed=31)
    for i in rng.choice(np.arange(so.groups[0].count), size=4, replace=False):
        mm = so.groups[0].meas[i].copy()
        mm[9:12] += rng.normal(0, 8.0, 3)
        mm[:9] = mm[:9].reshape(3, 3)[[1, 2, 0]].ravel()
        so.groups[0].meas[i] = mm
    so.save(os.path.join(HERE, "sphere_tiny_outliers.prob.bin"))
    bo = datasets.make("bal_tiny", seed=41)
    jdx = rng.choice(np.arange(bo.groups[0].count), size=6, replace=False)
    bo.groups[0].meas[jdx] += rng.normal(0, 60.0, (6, 2))
    bo.save(os.path.join(HERE, "bal_tiny_outliers.prob.bin"))
    for case in ("sphere_tiny_outliers", "bal_tiny_outliers"):
        for loss, nm in ((1, "tls"), (0, "gm")):
            subprocess.check_call([H, "gnc", os.path.join(HERE, f"{case}.prob.bin"), os.path.join(HERE, f"{case}.gnc_{nm}.bin"), str(loss)])
    # joint marginals of a few variable sets (Marginals::jointMarginalCovariance)
    sys.path.insert(0, os.path.dirname(HERE))
    import util as _util
    for case, sets in _util.JOINT_SETS.items():
        for i, vs in enumerate(sets):
            subprocess.check_call([H, "jointmarg", os.path.join(HERE, f"{case}.prob.bin"), os.path.join(HERE, f"{case}.joint{i}.bin")]
                                  + [str(v) for v in vs])
    # the reference's own end-to-end golden: tests/testGeneralSFMFactorB.cpp:44-63 (0.0199833 +- 1e-5)
    from gtsam_b200.problem import Problem
    with tempfile.TemporaryDirectory() as td:
        for mode, nm in ((0, "dubrovnik_3_7_unit"), (1, "dubrovnik_3_7_priors")):
            tmp = os.path.join(td, "p.bin")
            subprocess.check_call([H, "balfile", os.path.join(REF_DATA, "dubrovnik-3-7-pre.txt"), tmp, str(mode)])
            emit(nm, Problem.load(tmp), lm_iters=100)


if __name__ == "__main__":
    main()
