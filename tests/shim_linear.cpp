/*
 * shim_linear.cpp — TEST INFRASTRUCTURE: the GaussianFactorGraph level of the drop-in against the unmodified
 * reference, on real GTSAM objects.
 *
 *   shim_linear graph <problem.lin.bin>
 *       gtsam::GaussianFactorGraph::optimize(ordering, EliminatePreferCholesky)   (reference, CPU)
 *       gtsam_b200::optimizeOnDevice(gfg, ordering)                               (GPU through the C-ABI)
 *       + B200LinearSolver reused on a perturbed copy of the graph (same structure => numbers only)
 *   shim_linear bayestree <problem.lin.bin>    (no GPU needed) gtsam_b200::bayesTreeFromTables round trip on the reference's tree
 *   shim_linear hostpack <problem.lin.bin>     (no GPU needed)
 *       the shim's packing of the GaussianFactorGraph + the library's host symbolic phase vs the cliques of the
 *       reference's eliminateMultifrontal
 *   shim_linear pose2 <file.g2o> [maxit]
 *       BASELINE.json configs[0]: the Pose2 g2o graph + the example's prior (examples/Pose2SLAMExample_g2o.cpp:46-64);
 *       stock gtsam::LevenbergMarquardtOptimizer vs gtsam_b200::B200SolveLevenbergMarquardtOptimizer (solve() seam on
 *       the device, linearize / retract / error the reference's own), and the Gauss-Newton pair, as the example ships.
 * Prints one JSON line.  Built into oracle/_ref/shim_linear by gtsam_b200/shim/Makefile; run on the GPU box by
 * tests/test_gpu_shim_linear.py.
 */
#include "../oracle/linear_io.hpp"
#include "../gtsam_b200/shim/B200Optimizers.h"

#include <gtsam/geometry/Pose2.h>
#include <gtsam/linear/GaussianBayesTree.h>
#include <gtsam/linear/HessianFactor.h>
#include <gtsam/linear/linearExceptions.h>
#include <gtsam/slam/dataset.h>

#include <cmath>

using namespace gtsam;

static double relDiff(const VectorValues& a, const VectorValues& b) {
  double num = 0, den = 0;
  for (const auto& kv : b) {
    const Vector d = a.at(kv.first) - kv.second;
    num += d.squaredNorm();
    den += kv.second.squaredNorm();
  }
  return std::sqrt(num / std::max(den, 1e-300));
}

static int cmd_graph(const std::string& path) {
  linio::LinProb lp = linio::load(path);
  GaussianFactorGraph gfg = linio::build_graph(lp);
  Ordering ordering = linio::build_ordering(lp);
  int ref_status = 0, dev_status = 0;
  VectorValues ref, dev;
  try { ref = gfg.optimize(ordering, EliminatePreferCholesky); } catch (const IndeterminantLinearSystemException&) { ref_status = 1; }
  try { dev = gtsam_b200::optimizeOnDevice(gfg, ordering); } catch (const IndeterminantLinearSystemException&) { dev_status = 1; }
  double d0 = -1, d1 = -1, dbt = -1, dmarg = -1, dgrad = -1;
  int builds = 0, solves = 0;
  {   // GaussianFactorGraph::gradientAtZero() on the device vs the reference's (works on singular systems too)
    gtsam_b200::B200LinearSolver gs(ordering);
    dgrad = relDiff(gs.gradientAtZero(gfg), gfg.gradientAtZero());
  }
  long long launches = 0;
  if (!ref_status && !dev_status) {
    d0 = relDiff(dev, ref);
    // same structure, new numbers: only b200_linear_update + solve on the second call
    GaussianFactorGraph g2;
    std::mt19937 rng(5);
    std::normal_distribution<double> N(0, 1);
    for (const auto& f : gfg) {
      auto jf = std::dynamic_pointer_cast<JacobianFactor>(f);
      if (!jf) {   // a HessianFactor: scale its information a little (same structure, new numbers)
        auto hf = std::dynamic_pointer_cast<HessianFactor>(f);
        Matrix info = hf->info().selfadjointView();
        info *= 1.0 + 0.01 * std::fabs(N(rng));
        std::vector<DenseIndex> dims;
        for (auto it = hf->begin(); it != hf->end(); ++it) dims.push_back(hf->getDim(it));
        dims.push_back(1);
        g2.push_back(std::make_shared<HessianFactor>(hf->keys(), SymmetricBlockMatrix(dims, info)));
        continue;
      }
      Matrix Ab = jf->augmentedJacobianUnweighted();
      for (int c = 0; c < Ab.cols(); c++) for (int r = 0; r < Ab.rows(); r++) Ab(r, c) += 0.01 * N(rng);
      std::vector<std::pair<Key, Matrix>> terms;
      int col = 0;
      for (auto it = jf->begin(); it != jf->end(); ++it) { const int d = (int)jf->getDim(it); terms.emplace_back(*it, Ab.middleCols(col, d)); col += d; }
      g2.push_back(std::make_shared<JacobianFactor>(terms, Vector(Ab.col(col)), jf->get_model()));
    }
    gtsam_b200::B200LinearSolver solver(ordering);
    solver.optimize(gfg);
    const VectorValues dev2 = solver.optimize(g2);
    d1 = relDiff(dev2, g2.optimize(ordering, EliminatePreferCholesky));
    builds = solver.structureBuilds(); solves = solver.solves(); launches = solver.launchCount();
    // the device's elimination result as a real GaussianBayesTree
    auto bt = gtsam_b200::eliminateMultifrontalOnDevice(gfg, ordering);
    auto btr = gfg.eliminateMultifrontal(ordering, EliminatePreferCholesky);
    dbt = std::max(relDiff(bt->optimize(), ref), std::fabs(bt->logDeterminant() - btr->logDeterminant()) / std::max(1.0, std::fabs(btr->logDeterminant())));
    if (bt->size() != btr->size()) dbt = 1e300;
    // gtsam_b200::B200Marginals over the linear graph vs the marginals of the reference's Bayes tree
    gtsam_b200::B200Marginals marg(gfg, ordering);
    dmarg = 0;
    int taken = 0;
    for (Key k : gfg.keys()) {
      if (lp.var_dim[k] > 9 || taken++ >= 4) continue;
      const Matrix R = btr->marginalFactor(k, EliminatePreferCholesky)->information().inverse();
      dmarg = std::max(dmarg, (marg.marginalCovariance(k) - R).cwiseAbs().maxCoeff() / R.cwiseAbs().maxCoeff());
    }
  }
  printf("{\"ref_status\": %d, \"dev_status\": %d, \"delta_rel_diff\": %.6g, \"reuse_delta_rel_diff\": %.6g, \"bayes_tree_diff\": %.6g, \"marginals_diff\": %.6g, "
         "\"gradient_diff\": %.6g, \"structure_builds\": %d, \"solves\": %d, \"launches\": %lld}\n", ref_status, dev_status, d0, d1, dbt, dmarg, dgrad, builds, solves, launches);
  return 0;
}

/* host only: the shim's packing + the library's symbolic phase against the reference's Bayes tree */
static int cmd_hostpack(const std::string& path) {
  linio::LinProb lp = linio::load(path);
  GaussianFactorGraph gfg = linio::build_graph(lp);
  Ordering ordering = linio::build_ordering(lp);
  auto mine = gtsam_b200::symbolicOnHost(gfg, ordering);
  typedef std::pair<std::vector<Key>, std::vector<Key>> CS;
  std::vector<CS> a, b;
  for (auto& c : mine) { CS x(std::vector<Key>(c.first.begin(), c.first.end()), std::vector<Key>(c.second.begin(), c.second.end())); std::sort(x.second.begin(), x.second.end()); a.push_back(x); }
  GaussianFactorGraph sys = gfg;
  for (int64_t v = 0; v < lp.nvars; v++) {   // damping priors keep a singular fixture eliminable; they do not change the structure
    const int d = lp.var_dim[v];
    sys.push_back(std::make_shared<JacobianFactor>(Key(v), Matrix::Identity(d, d), Vector::Zero(d)));
  }
  auto bt = sys.eliminateMultifrontal(ordering, EliminatePreferCholesky);
  std::vector<GaussianBayesTree::sharedClique> stack(bt->roots().begin(), bt->roots().end());
  while (!stack.empty()) {
    auto c = stack.back(); stack.pop_back();
    auto cond = c->conditional();
    CS x(std::vector<Key>(cond->beginFrontals(), cond->endFrontals()), std::vector<Key>(cond->beginParents(), cond->endParents()));
    std::sort(x.second.begin(), x.second.end());
    b.push_back(x);
    for (auto& ch : c->children) stack.push_back(ch);
  }
  std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
  printf("{\"cliques\": %zu, \"reference_cliques\": %zu, \"equal\": %d}\n", a.size(), b.size(), (int)(a == b));
  return 0;
}

/* host only: gtsam_b200::bayesTreeFromTables on tables extracted from the REFERENCE's own Bayes tree must give a tree with
 * the same solution, determinant and clique count (round trip of what eliminateMultifrontalOnDevice does with the device's output) */
static int cmd_bayestree(const std::string& path) {
  linio::LinProb lp = linio::load(path);
  GaussianFactorGraph gfg = linio::build_graph(lp);
  for (int64_t v = 0; v < lp.nvars; v++) {   // keeps the singular fixture eliminable
    const int d = lp.var_dim[v];
    gfg.push_back(std::make_shared<JacobianFactor>(Key(v), Matrix::Identity(d, d), Vector::Zero(d)));
  }
  Ordering ordering = linio::build_ordering(lp);
  auto ref = gfg.eliminateMultifrontal(ordering, EliminatePreferCholesky);
  // flatten: children before parents
  std::vector<GaussianBayesTree::sharedClique> order;
  std::vector<GaussianBayesTree::sharedClique> stack(ref->roots().begin(), ref->roots().end());
  while (!stack.empty()) { auto c = stack.back(); stack.pop_back(); order.push_back(c); for (auto& ch : c->children) stack.push_back(ch); }
  std::reverse(order.begin(), order.end());
  std::map<const GaussianBayesTreeClique*, int64_t> index;
  for (size_t i = 0; i < order.size(); i++) index[order[i].get()] = (int64_t)i;
  std::vector<KeyVector> F, S;
  std::vector<int64_t> par;
  std::vector<Matrix> C;
  std::map<Key, int> dims;
  for (int64_t v = 0; v < lp.nvars; v++) dims[Key(v)] = lp.var_dim[v];
  for (auto& c : order) {
    auto cond = c->conditional();
    F.emplace_back(cond->beginFrontals(), cond->endFrontals());
    S.emplace_back(cond->beginParents(), cond->endParents());
    C.push_back(cond->augmentedJacobian());
    par.push_back(c->parent() ? index.at(c->parent().get()) : -1);
  }
  auto mine = gtsam_b200::bayesTreeFromTables(F, S, par, C, dims);
  const double d = relDiff(mine->optimize(), ref->optimize());
  printf("{\"cliques\": %zu, \"reference_cliques\": %zu, \"optimize_rel_diff\": %.6g, \"log_determinant_diff\": %.6g}\n", mine->size(), ref->size(), d,
         std::fabs(mine->logDeterminant() - ref->logDeterminant()));
  return 0;
}

template <class OPT>
static std::vector<double> run(OPT& opt, int maxit) {
  std::vector<double> errs{opt.error()};
  const auto& prm = opt.params();
  double currentError, newError = opt.error();
  do {
    currentError = newError;
    opt.iterate();
    newError = opt.error();
    errs.push_back(newError);
  } while ((int)opt.iterations() < maxit &&
           !checkConvergence(prm.relativeErrorTol, prm.absoluteErrorTol, prm.errorTol, currentError, newError) && std::isfinite(currentError));
  return errs;
}

static void printv(const char* name, const std::vector<double>& v) {
  printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); i++) printf("%s%.17g", i ? ", " : "", v[i]);
  printf("]");
}

static double valueDiff(const Values& a, const Values& b) {
  double m = 0;
  for (const auto& kv : b) m = std::max(m, kv.value.localCoordinates_(a.at(kv.key)).cwiseAbs().maxCoeff());
  return m;
}

static int cmd_pose2(const std::string& path, int maxit) {
  auto [graph, initial] = readG2o(path, false);
  graph->addPrior(0, Pose2(), noiseModel::Diagonal::Variances(Vector3(1e-6, 1e-6, 1e-8)));
  LevenbergMarquardtParams lp;
  lp.maxIterations = maxit;
  LevenbergMarquardtOptimizer ref(*graph, *initial, lp);
  gtsam_b200::B200SolveLevenbergMarquardtOptimizer dev(*graph, *initial, lp);
  const std::vector<double> e_ref = run(ref, maxit), e_dev = run(dev, maxit);
  GaussNewtonParams gp;
  gp.maxIterations = maxit;
  GaussNewtonOptimizer gref(*graph, *initial, gp);
  gtsam_b200::B200SolveGaussNewtonOptimizer gdev(*graph, *initial, gp);
  const std::vector<double> g_ref = run(gref, maxit), g_dev = run(gdev, maxit);
  printf("{");
  printv("lm_ref_errors", e_ref); printf(", "); printv("lm_dev_errors", e_dev); printf(", ");
  printv("gn_ref_errors", g_ref); printf(", "); printv("gn_dev_errors", g_dev);
  printf(", \"lm_value_diff\": %.6g, \"gn_value_diff\": %.6g, \"lm_ref_inner\": %d, \"lm_dev_inner\": %d, "
         "\"structure_builds\": %d, \"solves\": %d, \"launches\": %lld}\n",
         valueDiff(dev.values(), ref.values()), valueDiff(gdev.values(), gref.values()), ref.getInnerIterations(), dev.getInnerIterations(),
         dev.linearSolver().structureBuilds(), dev.linearSolver().solves(), dev.linearSolver().launchCount());
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 3 && std::string(argv[1]) == "graph") return cmd_graph(argv[2]);
  if (argc >= 3 && std::string(argv[1]) == "hostpack") return cmd_hostpack(argv[2]);
  if (argc >= 3 && std::string(argv[1]) == "bayestree") return cmd_bayestree(argv[2]);
  if (argc >= 3 && std::string(argv[1]) == "pose2") return cmd_pose2(argv[2], argc > 3 ? atoi(argv[3]) : 30);
  fprintf(stderr, "usage: shim_linear graph problem.lin.bin | pose2 file.g2o [maxit]\n");
  return 2;
}
