/*
 * shim_families.cpp — TEST INFRASTRUCTURE: factor families OUTSIDE the device-resident kinds, run through the
 * NonlinearOptimizer::solve() seam of the drop-in (gtsam_b200::B200SolveLevenbergMarquardtOptimizer: GTSAM linearizes
 * on the host, the GaussianFactorGraph — JacobianFactors of any arity, HessianFactors — is solved on the device).
 * These are the "rank 4" rows of SURVEY 8(f):
 *   sfm2   GeneralSFMFactor2<Cal3_S2> (pose, point AND calibration variables: ternary factors, gtsam/slam/GeneralSFMFactor.h:208)
 *   smart  SmartProjectionPoseFactor<Cal3_S2> in its default HESSIAN linearization mode (one RegularHessianFactor<6> over all
 *          the cameras of a point, gtsam/slam/SmartProjectionFactor.h / SmartFactorBase.h; timing/timeSFMBALsmart.cpp)
 *   expr   ExpressionFactor<Point2> — the autodiff BAL formulation (gtsam/nonlinear/ExpressionFactor.h, gtsam/slam/expressions.h;
 *          timing/timeSFMBALautodiff.cpp, examples/SFMExampleExpressions.cpp)
 * on a small synthetic scene (6 cameras on a ring, 24 points).
 *
 *   shim_families host   no GPU: the reference linearizes each graph; the shim packs the result and the library's host
 *                        symbolic phase must build the cliques of the reference's eliminateMultifrontal
 *   shim_families gpu    stock gtsam::LevenbergMarquardtOptimizer vs B200SolveLevenbergMarquardtOptimizer: error traces
 *   shim_families dumplin <dir>   the reference's linearizations as linear-problem files (golden inputs)
 * One JSON line.  Built into oracle/_ref/shim_families by gtsam_b200/shim/Makefile.
 */
#include "../gtsam_b200/shim/B200Optimizers.h"
#include "../oracle/linear_io.hpp"

#include <gtsam/geometry/Cal3_S2.h>
#include <gtsam/geometry/PinholeCamera.h>
#include <gtsam/inference/Symbol.h>
#include <gtsam/linear/GaussianBayesTree.h>
#include <gtsam/linear/HessianFactor.h>
#include <gtsam/linear/JacobianFactor.h>
#include <gtsam/nonlinear/ExpressionFactorGraph.h>
#include <gtsam/nonlinear/PriorFactor.h>
#include <gtsam/slam/GeneralSFMFactor.h>
#include <gtsam/slam/SmartProjectionPoseFactor.h>
#include <gtsam/slam/expressions.h>

#include <cmath>
#include <cstdio>
#include <random>
#include <string>

using namespace gtsam;
using symbol_shorthand::K;
using symbol_shorthand::L;
using symbol_shorthand::X;

struct Scene {
  std::vector<Pose3> poses;
  std::vector<Point3> points;
  Cal3_S2 cal{500, 500, 0, 320, 240};
  std::vector<std::vector<Point2>> z;   // z[j][i]: point j in camera i (every camera sees every point)
  Values initialPoses, initialPoints;
};

static Scene makeScene(int ncams = 6, int npoints = 24, unsigned seed = 11) {
  Scene s;
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> U(-2.0, 2.0);
  std::normal_distribution<double> N(0.0, 1.0);
  for (int i = 0; i < ncams; i++) {
    const double th = 2 * M_PI * i / ncams;
    const Point3 eye(15 * std::cos(th), 15 * std::sin(th), 1.5 * std::sin(3 * th));
    s.poses.push_back(PinholeBase::LookatPose(eye, Point3(0, 0, 0), Point3(0, 0, 1)));
  }
  for (int j = 0; j < npoints; j++) s.points.push_back(Point3(U(rng), U(rng), U(rng)));
  s.z.resize(npoints);
  for (int j = 0; j < npoints; j++)
    for (int i = 0; i < ncams; i++)
      s.z[j].push_back(PinholeCamera<Cal3_S2>(s.poses[i], s.cal).project(s.points[j]) + Point2(0.5 * N(rng), 0.5 * N(rng)));
  for (int i = 0; i < ncams; i++) {
    Vector6 d;
    for (int k = 0; k < 6; k++) d(k) = 0.01 * N(rng);
    s.initialPoses.insert(X(i), s.poses[i].retract(d));
  }
  for (int j = 0; j < npoints; j++) s.initialPoints.insert(L(j), Point3(s.points[j] + Point3(0.05 * N(rng), 0.05 * N(rng), 0.05 * N(rng))));
  return s;
}

struct Family {
  std::string name;
  NonlinearFactorGraph graph;
  Values initial;
};

static std::vector<Family> makeFamilies(const Scene& s) {
  const auto pix = noiseModel::Isotropic::Sigma(2, 1.0);
  const auto posePrior = noiseModel::Diagonal::Sigmas((Vector(6) << Vector3::Constant(0.05), Vector3::Constant(0.1)).finished());
  std::vector<Family> out;
  {  // GeneralSFMFactor2: calibration is a variable
    Family f;
    f.name = "sfm2";
    for (size_t j = 0; j < s.points.size(); j++)
      for (size_t i = 0; i < s.poses.size(); i++)
        f.graph.emplace_shared<GeneralSFMFactor2<Cal3_S2>>(s.z[j][i], pix, X(i), L(j), K(0));
    f.graph.addPrior(X(0), s.poses[0], posePrior);
    f.graph.addPrior(X(1), s.poses[1], posePrior);
    f.graph.addPrior(K(0), s.cal, noiseModel::Diagonal::Sigmas((Vector(5) << 5, 5, 0.01, 2, 2).finished()));
    f.initial.insert(s.initialPoses);
    f.initial.insert(s.initialPoints);
    f.initial.insert(K(0), Cal3_S2(505, 495, 0, 322, 238));
    out.push_back(f);
  }
  {  // smart factors: the points are not variables
    Family f;
    f.name = "smart";
    auto Kp = std::make_shared<Cal3_S2>(s.cal);
    for (size_t j = 0; j < s.points.size(); j++) {
      auto sf = std::make_shared<SmartProjectionPoseFactor<Cal3_S2>>(pix, Kp);
      for (size_t i = 0; i < s.poses.size(); i++) sf->add(s.z[j][i], X(i));
      f.graph.push_back(sf);
    }
    f.graph.addPrior(X(0), s.poses[0], posePrior);
    f.graph.addPrior(X(1), s.poses[1], posePrior);
    f.initial.insert(s.initialPoses);
    out.push_back(f);
  }
  {  // expressions (autodiff)
    Family f;
    f.name = "expr";
    ExpressionFactorGraph g;
    Cal3_S2_ cK(s.cal);
    for (size_t j = 0; j < s.points.size(); j++)
      for (size_t i = 0; i < s.poses.size(); i++) {
        Pose3_ x(X(i));
        Point3_ p(L(j));
        g.addExpressionFactor(uncalibrate(cK, project(transformTo(x, p))), s.z[j][i], pix);
      }
    g.addExpressionFactor(Pose3_(X(0)), s.poses[0], posePrior);
    g.addExpressionFactor(Pose3_(X(1)), s.poses[1], posePrior);
    f.graph = g;
    f.initial.insert(s.initialPoses);
    f.initial.insert(s.initialPoints);
    out.push_back(f);
  }
  return out;
}

typedef std::pair<std::vector<Key>, std::vector<Key>> CS;
static std::vector<CS> referenceCliques(const GaussianFactorGraph& gfg, const Ordering& ordering) {
  std::vector<CS> b;
  auto bt = gfg.eliminateMultifrontal(ordering, EliminatePreferCholesky);
  std::vector<GaussianBayesTree::sharedClique> stack(bt->roots().begin(), bt->roots().end());
  while (!stack.empty()) {
    auto c = stack.back(); stack.pop_back();
    auto cond = c->conditional();
    CS x(std::vector<Key>(cond->beginFrontals(), cond->endFrontals()), std::vector<Key>(cond->beginParents(), cond->endParents()));
    std::sort(x.second.begin(), x.second.end());
    b.push_back(x);
    for (auto& ch : c->children) stack.push_back(ch);
  }
  std::sort(b.begin(), b.end());
  return b;
}

static int cmd_host() {
  const Scene s = makeScene();
  printf("{");
  bool first = true;
  for (auto& f : makeFamilies(s)) {
    auto lin = f.graph.linearize(f.initial);
    const Ordering ordering = Ordering::Colamd(*lin);
    int nj = 0, nh = 0, other = 0;
    size_t maxArity = 0;
    for (auto& g : *lin) {
      if (std::dynamic_pointer_cast<JacobianFactor>(g)) nj++;
      else if (std::dynamic_pointer_cast<HessianFactor>(g)) nh++;
      else other++;
      maxArity = std::max(maxArity, g->size());
    }
    std::vector<CS> a;
    for (auto& c : gtsam_b200::symbolicOnHost(*lin, ordering)) {
      CS x(std::vector<Key>(c.first.begin(), c.first.end()), std::vector<Key>(c.second.begin(), c.second.end()));
      std::sort(x.second.begin(), x.second.end());
      a.push_back(x);
    }
    std::sort(a.begin(), a.end());
    const std::vector<CS> b = referenceCliques(*lin, ordering);
    printf("%s\"%s\": {\"jacobian\": %d, \"hessian\": %d, \"other\": %d, \"max_arity\": %zu, \"cliques\": %zu, \"cliques_equal\": %d}",
           first ? "" : ", ", f.name.c_str(), nj, nh, other, maxArity, a.size(), (int)(a == b));
    first = false;
  }
  printf("}\n");
  return 0;
}

/* the reference's linearization of every family as a linear-problem file (tests/golden/lin_family_<name>.lin.bin) */
static int cmd_dumplin(const std::string& dir) {
  const Scene s = makeScene();
  for (auto& f : makeFamilies(s)) {
    auto lin = f.graph.linearize(f.initial);
    const Ordering ordering = Ordering::Colamd(*lin);
    linio::LinProb lp;
    if (!linio::from_graph(*lin, ordering, &lp)) { fprintf(stderr, "%s: unsupported Gaussian factor\n", f.name.c_str()); return 3; }
    linio::save(lp, dir + "/lin_family_" + f.name + ".lin.bin");
    printf("%s: %ld variables, %ld factors, %zu groups\n", f.name.c_str(), (long)lp.nvars, (long)lp.nfactors(), lp.groups.size());
  }
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

static int cmd_gpu() {
  const Scene s = makeScene();
  printf("{");
  bool first = true;
  for (auto& f : makeFamilies(s)) {
    LevenbergMarquardtParams lp;
    lp.maxIterations = 20;
    LevenbergMarquardtOptimizer ref(f.graph, f.initial, lp);
    gtsam_b200::B200SolveLevenbergMarquardtOptimizer dev(f.graph, f.initial, lp);
    const std::vector<double> er = run(ref, 20), ed = run(dev, 20);
    double worst = er.size() == ed.size() ? 0.0 : 1e300;
    for (size_t i = 0; i < std::min(er.size(), ed.size()); i++) worst = std::max(worst, std::fabs(er[i] - ed[i]) / std::max(1e-12, std::fabs(er[i])));
    double vdiff = 0;
    for (const auto& kv : ref.values()) vdiff = std::max(vdiff, kv.value.localCoordinates_(dev.values().at(kv.key)).cwiseAbs().maxCoeff());
    printf("%s\"%s\": {\"iterations\": %zu, \"worst_error_rel_diff\": %.6g, \"value_diff\": %.6g, \"final_error\": %.12g, \"solves\": %d, \"builds\": %d, \"launches\": %lld}",
           first ? "" : ", ", f.name.c_str(), er.size() - 1, worst, vdiff, ed.back(), dev.linearSolver().solves(), dev.linearSolver().structureBuilds(),
           dev.linearSolver().launchCount());
    first = false;
  }
  printf("}\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "host") return cmd_host();
  if (argc >= 2 && std::string(argv[1]) == "gpu") return cmd_gpu();
  if (argc >= 3 && std::string(argv[1]) == "dumplin") return cmd_dumplin(argv[2]);
  fprintf(stderr, "usage: shim_families host | gpu | dumplin <dir>\n");
  return 2;
}
