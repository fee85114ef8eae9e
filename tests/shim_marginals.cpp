/*
 * shim_marginals.cpp — TEST INFRASTRUCTURE: gtsam::Marginals (unmodified reference, CPU) against the drop-in
 * gtsam_b200::B200Marginals, and gtsam::DoglegOptimizer against gtsam_b200::B200DoglegOptimizer (GPU through the
 * C-ABI), on the same NonlinearFactorGraph / Values / Ordering.
 * Prints one JSON line.  Built into oracle/_ref/shim_marginals by gtsam_b200/shim/Makefile; run on the GPU box by
 * tests/test_gpu_shim.py.
 */
#include "../oracle/problem_io.hpp"
#include "../gtsam_b200/shim/B200Optimizers.h"

#include <gtsam/nonlinear/GncOptimizer.h>

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: shim_marginals problem.bin [run_gnc]\n"); return 2; }
  Prob p = load(argv[1]);
  Built b = build(p);
  Marginals ref(b.graph, b.values, b.ordering, Marginals::CHOLESKY);
  gtsam_b200::B200Marginals dev(b.graph, b.values, b.ordering);
  double worst = 0, worst_info = 0, worst_joint = 0;
  for (int64_t v = 0; v < p.nvars; v++) {
    const Matrix R = ref.marginalCovariance(Key(v)), S = dev.marginalCovariance(Key(v));
    worst = std::max(worst, (R - S).cwiseAbs().maxCoeff() / R.cwiseAbs().maxCoeff());
  }
  {
    const Matrix R = ref.marginalInformation(Key(0)), S = dev.marginalInformation(Key(0));
    worst_info = (R - S).cwiseAbs().maxCoeff() / R.cwiseAbs().maxCoeff();
  }
  {
    KeyVector keys{Key(p.nvars - 1), Key(0), Key(p.nvars / 2)};   // deliberately unsorted
    const Matrix R = ref.jointMarginalCovariance(keys).fullMatrix(), S = dev.jointMarginalCovariance(keys);
    worst_joint = (R - S).cwiseAbs().maxCoeff() / R.cwiseAbs().maxCoeff();
  }
  // Dogleg drop-in against the stock DoglegOptimizer: errors and trust-region radii of 5 iterations
  double dl_err = 0, dl_delta = 0, dl_values = 0;
  {
    DoglegParams dp;
    dp.ordering = b.ordering;
    DoglegOptimizer sref(b.graph, b.values, dp);
    gtsam_b200::B200DoglegOptimizer sdev(b.graph, b.values, dp);
    for (int i = 0; i < 5; i++) {
      sref.iterate(); sdev.iterate();
      dl_err = std::max(dl_err, std::abs(sref.error() - sdev.error()) / std::max(1.0, std::abs(sref.error())));
      dl_delta = std::max(dl_delta, std::abs(sref.getDelta() - sdev.getDelta()) / std::max(1e-12, std::abs(sref.getDelta())));
    }
    for (const auto& kv : sref.values()) {
      Vector d = kv.value.localCoordinates_(sdev.values().at(kv.key));
      dl_values = std::max(dl_values, d.cwiseAbs().maxCoeff());
    }
    if (sdev.iterations() != 5) dl_err = 1;
  }
  // the reference's own GncOptimizer template with the device LM plugged in through the params type
  double gnc_w = -1, gnc_v = -1;
  if (argc > 2 && atoi(argv[2])) {
    LevenbergMarquardtParams lmp;
    lmp.ordering = b.ordering;
    GncParams<LevenbergMarquardtParams> gp(lmp);
    GncOptimizer<GncParams<LevenbergMarquardtParams>> gref(b.graph, b.values, gp);
    const Values rref = gref.optimize();
    GncParams<gtsam_b200::B200LevenbergMarquardtParams> gd{gtsam_b200::B200LevenbergMarquardtParams(lmp)};
    GncOptimizer<GncParams<gtsam_b200::B200LevenbergMarquardtParams>> gdev(b.graph, b.values, gd);
    const Values rdev = gdev.optimize();
    gnc_w = (gref.getWeights() - gdev.getWeights()).cwiseAbs().maxCoeff();
    gnc_v = 0;
    for (const auto& kv : rref) gnc_v = std::max(gnc_v, kv.value.localCoordinates_(rdev.at(kv.key)).cwiseAbs().maxCoeff());
  }
  printf("{\"gnc_weights\": %.6g, \"gnc_values\": %.6g, ", gnc_w, gnc_v);
  printf("\"dogleg_error\": %.6g, \"dogleg_delta\": %.6g, \"dogleg_values\": %.6g, ", dl_err, dl_delta, dl_values);
  printf("\"worst_cov\": %.6g, \"worst_info\": %.6g, \"worst_joint\": %.6g, \"variables\": %lld}\n", worst, worst_info, worst_joint,
         (long long)p.nvars);
  return 0;
}
