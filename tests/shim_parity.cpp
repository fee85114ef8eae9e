/*
 * shim_parity.cpp — TEST INFRASTRUCTURE: runs the stock gtsam::LevenbergMarquardtOptimizer
 * (unmodified reference, CPU) and the drop-in gtsam_b200::B200LevenbergMarquardtOptimizer
 * (GPU through the C-ABI) side by side on the same NonlinearFactorGraph / Values /
 * Ordering and prints one JSON line with the per-iteration traces of both.
 * Built into oracle/_ref/shim_parity by gtsam_b200/shim/Makefile; run on the GPU box
 * by tests/test_gpu_shim.py.
 */
#include "../oracle/problem_io.hpp"
#include "../gtsam_b200/shim/B200Optimizers.h"
#include <unistd.h>

#include <chrono>

template <class OPT>
static void trace(OPT& opt, int maxit, std::vector<double>& errs, std::vector<double>& lams, std::vector<int>& inner, double& secs) {
  const auto& prm = opt.params();
  errs.push_back(opt.error()); lams.push_back(opt.lambda()); inner.push_back(0);
  double currentError, newError = opt.error();
  auto t0 = std::chrono::high_resolution_clock::now();
  do {
    currentError = newError;
    opt.iterate();
    newError = opt.error();
    errs.push_back(newError); lams.push_back(opt.lambda()); inner.push_back(opt.getInnerIterations());
  } while ((int)opt.iterations() < maxit &&
           !checkConvergence(prm.relativeErrorTol, prm.absoluteErrorTol, prm.errorTol, currentError, newError) &&
           std::isfinite(currentError));
  secs = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
}

static void printv(const char* name, const std::vector<double>& v) {
  printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); i++) printf("%s%.17g", i ? ", " : "", v[i]);
  printf("]");
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: shim_parity problem.bin [maxit] [ceres] [skip_reference]\n"); return 2; }
  const int maxit = argc > 2 ? atoi(argv[2]) : 30;
  const bool ceres = argc > 3 && atoi(argv[3]);
  const bool skipRef = argc > 4 && atoi(argv[4]);
  Prob p = load(argv[1]);
  Built b = build(p);
  LevenbergMarquardtParams params = ceres ? LevenbergMarquardtParams::CeresDefaults() : LevenbergMarquardtParams::LegacyDefaults();
  params.ordering = b.ordering;
  params.maxIterations = maxit;
  std::vector<double> e_ref, l_ref, e_dev, l_dev;
  std::vector<int> i_ref, i_dev;
  double t_ref = 0, t_dev = 0, maxdiff = 0;
  // sharded mode: shim_parity problem.bin maxit ceres 0 <world> <rank> <uid file> [device]: one process per rank, the
  // communicator id travels through a file (rank 0 writes it); every rank runs the same B200LevenbergMarquardtOptimizer
  // constructor + the communicator and compares with the stock optimizer on the same GTSAM objects
  if (argc > 7 && atoi(argv[5]) > 1) {
    gtsam_b200::B200Communicator comm;
    comm.world = atoi(argv[5]); comm.rank = atoi(argv[6]); comm.device = argc > 8 ? atoi(argv[8]) : comm.rank;
    const std::string path = argv[7], tmp = path + ".tmp";
    if (comm.rank == 0) {
      comm.uniqueId = gtsam_b200::B200Communicator::newUniqueId();
      FILE* fh = fopen(tmp.c_str(), "wb");
      fwrite(comm.uniqueId.data(), 1, 128, fh); fclose(fh);
      rename(tmp.c_str(), path.c_str());
    } else {
      FILE* fh = nullptr;
      for (int tries = 0; tries < 60000 && !(fh = fopen(path.c_str(), "rb")); tries++) usleep(1000);
      if (!fh || fread(comm.uniqueId.data(), 1, 128, fh) != 128) { fprintf(stderr, "no communicator id\n"); return 3; }
      fclose(fh);
    }
    gtsam_b200::B200LevenbergMarquardtOptimizer sh(b.graph, b.values, b.ordering, params, comm);
    trace(sh, maxit, e_dev, l_dev, i_dev, t_dev);
    Values shValues = sh.values();      // the FULL estimate on every rank
    LevenbergMarquardtOptimizer ref(b.graph, b.values, params);
    trace(ref, maxit, e_ref, l_ref, i_ref, t_ref);
    for (const auto& kv : ref.values()) {
      Vector d = kv.value.localCoordinates_(shValues.at(kv.key));
      maxdiff = std::max(maxdiff, d.cwiseAbs().maxCoeff());
    }
    printf("{\"world\": %d, \"rank\": %d, ", comm.world, comm.rank);
    printv("dev_errors", e_dev); printf(", "); printv("dev_lambdas", l_dev); printf(", ");
    printv("ref_errors", e_ref); printf(", "); printv("ref_lambdas", l_ref);
    printf(", \"dev_inner\": [");
    for (size_t i = 0; i < i_dev.size(); i++) printf("%s%d", i ? ", " : "", i_dev[i]);
    printf("], \"ref_inner\": [");
    for (size_t i = 0; i < i_ref.size(); i++) printf("%s%d", i ? ", " : "", i_ref[i]);
    printf("], \"max_value_diff\": %.6g, \"launches\": %lld}\n", maxdiff, sh.launchCount());
    return 0;
  }
  gtsam_b200::B200LevenbergMarquardtOptimizer dev(b.graph, b.values, params);
  trace(dev, maxit, e_dev, l_dev, i_dev, t_dev);
  // the hook / public accessors keep working through the base class
  Values devValues = dev.values();
  if (!skipRef) {
    LevenbergMarquardtOptimizer ref(b.graph, b.values, params);
    trace(ref, maxit, e_ref, l_ref, i_ref, t_ref);
    for (const auto& kv : ref.values()) {
      Vector d = kv.value.localCoordinates_(devValues.at(kv.key));
      maxdiff = std::max(maxdiff, d.cwiseAbs().maxCoeff());
    }
  }
  // linearize() seam: device Jacobians as a GaussianFactorGraph vs the reference's
  gtsam_b200::B200LevenbergMarquardtOptimizer dev2(b.graph, b.values, params);
  auto lin_dev = dev2.linearize();
  auto lin_ref = b.graph.linearize(b.values);
  double jdiff = 0;
  for (size_t i = 0; i < lin_ref->size(); i++) {
    Matrix A = std::dynamic_pointer_cast<JacobianFactor>((*lin_ref)[i])->augmentedJacobian();
    Matrix B = std::dynamic_pointer_cast<JacobianFactor>((*lin_dev)[i])->augmentedJacobian();
    jdiff = std::max(jdiff, (A - B).cwiseAbs().maxCoeff() / std::max(1.0, A.cwiseAbs().maxCoeff()));
  }
  // optimize() through the unmodified base-class loop, with an iterationHook
  int hooks = 0;
  LevenbergMarquardtParams p2 = params;
  p2.iterationHook = [&](size_t, double, double) { hooks++; };
  gtsam_b200::B200LevenbergMarquardtOptimizer dev3(b.graph, b.values, p2);
  Values res = dev3.optimize();
  printf("{");
  printv("dev_errors", e_dev); printf(", "); printv("dev_lambdas", l_dev); printf(", ");
  printv("ref_errors", e_ref); printf(", "); printv("ref_lambdas", l_ref);
  printf(", \"dev_inner\": [");
  for (size_t i = 0; i < i_dev.size(); i++) printf("%s%d", i ? ", " : "", i_dev[i]);
  printf("], \"ref_inner\": [");
  for (size_t i = 0; i < i_ref.size(); i++) printf("%s%d", i ? ", " : "", i_ref[i]);
  printf("], \"max_value_diff\": %.6g, \"linearize_max_rel_diff\": %.6g, \"dev_seconds\": %.6f, \"ref_seconds\": %.6f, "
         "\"optimize_error\": %.17g, \"optimize_iterations\": %d, \"hook_calls\": %d, \"launches\": %lld}\n",
         maxdiff, jdiff, t_dev, t_ref, dev3.error(), (int)dev3.iterations(), hooks, dev.launchCount());
  return 0;
}
