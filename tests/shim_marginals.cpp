/*
 * shim_marginals.cpp — TEST INFRASTRUCTURE: gtsam::Marginals (unmodified reference, CPU) against the drop-in
 * gtsam_b200::B200Marginals (GPU through the C-ABI) on the same NonlinearFactorGraph / Values / Ordering.
 * Prints one JSON line.  Built into oracle/_ref/shim_marginals by gtsam_b200/shim/Makefile; run on the GPU box by
 * tests/test_gpu_shim.py.
 */
#include "../oracle/problem_io.hpp"
#include "../gtsam_b200/shim/B200Optimizers.h"

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: shim_marginals problem.bin\n"); return 2; }
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
  printf("{\"worst_cov\": %.6g, \"worst_info\": %.6g, \"worst_joint\": %.6g, \"variables\": %lld}\n", worst, worst_info, worst_joint,
         (long long)p.nvars);
  return 0;
}
