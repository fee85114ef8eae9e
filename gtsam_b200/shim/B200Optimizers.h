/**
 * B200Optimizers.h — the GTSAM-side drop-in: subclasses of the reference's own
 * optimizers whose virtual seams run on the B200 through the C-ABI
 * (include/gtsam_b200.h).  User code changes one type name:
 *
 *     gtsam::LevenbergMarquardtOptimizer lm(graph, initial, params);          // before
 *     gtsam_b200::B200LevenbergMarquardtOptimizer lm(graph, initial, params);  // after
 *     Values result = lm.optimize();
 *
 * Seams overridden (reference file:line):
 *   - NonlinearOptimizer::iterate()            gtsam/nonlinear/NonlinearOptimizer.h:136
 *     (LevenbergMarquardtOptimizer.cpp:273-308, GaussNewtonOptimizer.cpp:44-67):
 *     linearize + damped multifrontal solve + retract + error stay on the device;
 *   - LevenbergMarquardtOptimizer::linearize() gtsam/nonlinear/LevenbergMarquardtOptimizer.h:112-113:
 *     returns the device linearization as whitened JacobianFactors (debug / parity).
 * params(), lambda(), error(), values(), iterations(), getInnerIterations(),
 * optimize(), iterationHook and verbosity behave as in the reference because
 * the unmodified base-class loop (NonlinearOptimizer::defaultOptimize) drives iterate().
 *
 * Supported factors (anything else => std::invalid_argument, there is no CPU fallback):
 * BetweenFactor<Pose3|Pose2>, PriorFactor<Pose3|Pose2|Point3|PinholeCamera<Cal3Bundler>>,
 * GenericProjectionFactor<Pose3,Point3,Cal3_S2> (with or without body_P_sensor),
 * GeneralSFMFactor<PinholeCamera<Cal3Bundler>,Point3>; noise models Unit,
 * Isotropic, Diagonal, Gaussian, and noiseModel::Robust (Huber / Cauchy / Tukey / Fair) around any
 * of them (Constrained => std::invalid_argument).
 */
#pragma once
#include <gtsam/nonlinear/DoglegOptimizer.h>
#include <gtsam/nonlinear/GaussNewtonOptimizer.h>
#include <gtsam/nonlinear/LevenbergMarquardtOptimizer.h>
#include <gtsam/nonlinear/Marginals.h>

#include <gtsam/linear/GaussianBayesTree.h>

#include <array>
#include <map>
#include <memory>
#include <vector>

namespace gtsam_b200 {

/// One rank of a sharded solve: one process per GPU, `world` of them.  Rank 0 draws the id
/// (B200Communicator::newUniqueId) and ships the 128 bytes to the other ranks by whatever the application has (MPI, a
/// file, a socket); every rank then constructs the same optimizer on the SAME graph, Values and Ordering.  The library
/// splits the junction tree (subtrees by rank, the top distributed by owner), every rank takes the same LM decisions and
/// values() holds the full estimate on every rank.
struct B200Communicator {
  int rank = 0, world = 1, device = 0;
  std::array<char, 128> uniqueId{};
  static std::array<char, 128> newUniqueId();
};

struct DeviceState;  // packed problem + C-ABI handles
struct LinearState;  // packed JacobianFactor / HessianFactor groups + C-ABI handles

class B200LevenbergMarquardtOptimizer : public gtsam::LevenbergMarquardtOptimizer {
 public:
  B200LevenbergMarquardtOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                                  const gtsam::LevenbergMarquardtParams& params = gtsam::LevenbergMarquardtParams());
  B200LevenbergMarquardtOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                                  const gtsam::Ordering& ordering,
                                  const gtsam::LevenbergMarquardtParams& params = gtsam::LevenbergMarquardtParams());
  /// Sharded over the ranks of `comm` (same constructor arguments as gtsam/nonlinear/LevenbergMarquardtOptimizer.h:59-72
  /// plus the communicator): iterate() / optimize() / values() / error() / lambda() behave as on one GPU.
  B200LevenbergMarquardtOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                                  const gtsam::Ordering& ordering, const gtsam::LevenbergMarquardtParams& params,
                                  const B200Communicator& comm);
  ~B200LevenbergMarquardtOptimizer() override;

  /// One LM iteration on the device.  Returns nullptr: the linear graph stays in HBM
  /// (call linearize() to materialise it on the host).
  gtsam::GaussianFactorGraph::shared_ptr iterate() override;
  gtsam::GaussianFactorGraph::shared_ptr linearize() const override;
  /// kernels launched so far (evidence that the device path ran)
  long long launchCount() const;
  /// true: the "FP32 linearize + FP64 solve" mode of BASELINE configs[4] (whitened Jacobians kept as floats,
  /// b200_set_jacobian_precision); false (default): the reference's FP64 arithmetic throughout
  void setJacobianFp32(bool on);

 private:
  void init(const B200Communicator* comm = nullptr);
  std::shared_ptr<DeviceState> dev_;
};

class B200GaussNewtonOptimizer : public gtsam::GaussNewtonOptimizer {
 public:
  B200GaussNewtonOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                           const gtsam::GaussNewtonParams& params = gtsam::GaussNewtonParams());
  B200GaussNewtonOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                           const gtsam::Ordering& ordering);
  ~B200GaussNewtonOptimizer() override;
  gtsam::GaussianFactorGraph::shared_ptr iterate() override;

 private:
  void init();
  std::shared_ptr<DeviceState> dev_;
};

/// Parameter type that names the device LM as its optimizer, so the reference's own templates pick it up:
///     gtsam::GncOptimizer<gtsam::GncParams<gtsam_b200::B200LevenbergMarquardtParams>> gnc(graph, initial, params);
/// runs graduated non-convexity (gtsam/nonlinear/GncOptimizer.h:184-268) with every weighted LM solve on the B200
/// (BaseOptimizer = GncParameters::OptimizerType, GncOptimizer.h:47, GncParams.h:44).
struct B200LevenbergMarquardtParams : public gtsam::LevenbergMarquardtParams {
  typedef B200LevenbergMarquardtOptimizer OptimizerType;
  B200LevenbergMarquardtParams() = default;
  B200LevenbergMarquardtParams(const gtsam::LevenbergMarquardtParams& p) : gtsam::LevenbergMarquardtParams(p) {}
};

/// Drop-in for gtsam::DoglegOptimizer (gtsam/nonlinear/DoglegOptimizer.h:63-128): same constructors, params(),
/// getDelta(), iterate(), optimize().  It derives from NonlinearOptimizer rather than from DoglegOptimizer because the
/// reference keeps its DoglegState private to DoglegOptimizer.cpp (getDelta() is non-virtual and casts to it); the
/// unmodified NonlinearOptimizer::defaultOptimize() drives iterate() exactly as it drives the reference's.
class B200DoglegOptimizer : public gtsam::NonlinearOptimizer {
 public:
  B200DoglegOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                      const gtsam::DoglegParams& params = gtsam::DoglegParams());
  B200DoglegOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                      const gtsam::Ordering& ordering);
  ~B200DoglegOptimizer() override;
  /// One dogleg iteration on the device (DoglegOptimizer.cpp:84-121); returns nullptr (the linear graph stays in HBM).
  gtsam::GaussianFactorGraph::shared_ptr iterate() override;
  const gtsam::DoglegParams& params() const { return params_; }
  /// current trust-region radius (DoglegOptimizer::getDelta)
  double getDelta() const;

 protected:
  const gtsam::NonlinearOptimizerParams& _params() const override { return params_; }
  gtsam::DoglegParams params_;

 private:
  void init();
  std::shared_ptr<DeviceState> dev_;
  void* dl_ = nullptr;   // b200_dl*
};

/// Drop-in for gtsam::Marginals (gtsam/nonlinear/Marginals.h:31-100, CHOLESKY factorization): covariances of the
/// graph linearised at `solution`, from the multifrontal factor kept on the device (one linearize + solve at
/// construction of the first query; each query walks the clique path(s) from the variable(s) to the root).
class B200Marginals {
 public:
  B200Marginals(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& solution, const gtsam::Ordering& ordering);
  /// ordering = Ordering::Colamd(graph), as gtsam::Marginals(graph, solution) computes it (Marginals.cpp:30-36)
  B200Marginals(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& solution);
  /// the linear constructors of gtsam::Marginals (Marginals.h:60-76): an already linear graph of Jacobian / Hessian factors
  B200Marginals(const gtsam::GaussianFactorGraph& graph, const gtsam::Ordering& ordering);
  explicit B200Marginals(const gtsam::GaussianFactorGraph& graph);
  /// Marginals::marginalCovariance, gtsam/nonlinear/Marginals.cpp:118-126
  gtsam::Matrix marginalCovariance(gtsam::Key variable) const;
  /// Marginals::marginalInformation, gtsam/nonlinear/Marginals.cpp:128-154
  gtsam::Matrix marginalInformation(gtsam::Key variable) const;
  /// JointMarginal::fullMatrix() of Marginals::jointMarginalCovariance: blocks in sorted-key order
  /// (gtsam::JointMarginal's constructor is private to gtsam::Marginals, hence the plain matrix)
  gtsam::Matrix jointMarginalCovariance(const gtsam::KeyVector& variables) const;

 private:
  std::shared_ptr<DeviceState> dev_;     // nonlinear graph + Values ...
  std::shared_ptr<LinearState> lin_;     // ... or a GaussianFactorGraph
};

/// GaussianFactorGraph::optimize(ordering, EliminatePreferCholesky) (gtsam/linear/GaussianFactorGraph.cpp:316-319) on the
/// device for ANY graph of JacobianFactors (any arity / block widths, Unit or Diagonal models) and HessianFactors: the
/// "GaussianFactorGraph::optimize-level entry" of SURVEY 8b.  Constrained models and other GaussianFactor types =>
/// std::invalid_argument (no CPU fallback); a singular system => IndeterminantLinearSystemException as in the reference.
gtsam::VectorValues optimizeOnDevice(const gtsam::GaussianFactorGraph& gfg, const gtsam::Ordering& ordering);

/// Host-only (needs no GPU): packs `gfg` exactly as optimizeOnDevice does and runs the library's symbolic phase;
/// returns (frontal keys, separator keys) of every clique in elimination order — the cliques the reference's
/// eliminateMultifrontal(ordering) builds.  For inspection and CPU-side tests of the packing.
std::vector<std::pair<gtsam::KeyVector, gtsam::KeyVector>> symbolicOnHost(const gtsam::GaussianFactorGraph& gfg,
                                                                           const gtsam::Ordering& ordering);

/// GaussianFactorGraph::eliminateMultifrontal(ordering, EliminatePreferCholesky)
/// (gtsam/inference/EliminateableFactorGraph-inst.h:123-146) on the device: the junction tree is eliminated there and the
/// result comes back as a real gtsam::GaussianBayesTree (one GaussianConditional [R S d] per clique), so everything the
/// reference offers on a Bayes tree (optimize(), determinant(), marginalFactor(), ...) works on it.
gtsam::GaussianBayesTree::shared_ptr eliminateMultifrontalOnDevice(const gtsam::GaussianFactorGraph& gfg, const gtsam::Ordering& ordering);

/// Host-only helper (needs no GPU): the GaussianBayesTree of flat clique tables — frontal / separator keys and parent
/// index of every clique (children before parents, -1 for roots) and its conditional [R S d] as an f x (f+s+1) matrix with
/// the columns in frontals-then-separators order.  It is what eliminateMultifrontalOnDevice calls on the device's output.
gtsam::GaussianBayesTree::shared_ptr bayesTreeFromTables(const std::vector<gtsam::KeyVector>& frontals,
                                                         const std::vector<gtsam::KeyVector>& separators,
                                                         const std::vector<int64_t>& parent, const std::vector<gtsam::Matrix>& conditionals,
                                                         const std::map<gtsam::Key, int>& dims);

/// The same, keeping the device problem: successive graphs with the SAME structure (the next linearization of one
/// nonlinear graph, the next lambda of LM's damped system) only re-upload numbers (b200_linear_update); the symbolic
/// phase, which the reference repeats inside every optimize() (VariableIndex + elimination tree + junction tree,
/// gtsam/inference/EliminateableFactorGraph-inst.h:123-146), runs once.
class B200LinearSolver {
 public:
  explicit B200LinearSolver(const gtsam::Ordering& ordering);
  ~B200LinearSolver();
  /// solve `gfg` (re-packs only if its structure differs from the previous call's)
  gtsam::VectorValues optimize(const gtsam::GaussianFactorGraph& gfg);
  /// GaussianFactorGraph::gradientAtZero() (gtsam/linear/GaussianFactorGraph.cpp:369-378) of `gfg` on the device
  gtsam::VectorValues gradientAtZero(const gtsam::GaussianFactorGraph& gfg);
  /// how many times the structure was (re)built / how many solves reused it
  int structureBuilds() const;
  int solves() const;
  long long launchCount() const;

 private:
  gtsam::Ordering ordering_;
  std::shared_ptr<LinearState> st_;
};

/// LevenbergMarquardtOptimizer whose linear-solve seam
///     virtual VectorValues NonlinearOptimizer::solve(const GaussianFactorGraph&, const NonlinearOptimizerParams&) const
/// (gtsam/nonlinear/NonlinearOptimizer.h:128-130, called by tryLambda at LevenbergMarquardtOptimizer.cpp:156) runs on the
/// device.  linearize(), the damped system, retract and error stay the reference's own host code, so this variant
/// accepts ANY factor type GTSAM can linearize to Jacobian / Hessian factors (the Pose2 graph of BASELINE configs[0],
/// GeneralSFMFactor2, SmartProjectionPoseFactor in HESSIAN mode, ExpressionFactors: tests/shim_families.cpp); use
/// B200LevenbergMarquardtOptimizer when all factors are of the device-resident kinds.
/// Requires linearSolverType MULTIFRONTAL_CHOLESKY (the default); anything else => std::invalid_argument.
class B200SolveLevenbergMarquardtOptimizer : public gtsam::LevenbergMarquardtOptimizer {
 public:
  B200SolveLevenbergMarquardtOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                                       const gtsam::LevenbergMarquardtParams& params = gtsam::LevenbergMarquardtParams());
  B200SolveLevenbergMarquardtOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                                       const gtsam::Ordering& ordering,
                                       const gtsam::LevenbergMarquardtParams& params = gtsam::LevenbergMarquardtParams());
  gtsam::VectorValues solve(const gtsam::GaussianFactorGraph& gfg, const gtsam::NonlinearOptimizerParams& params) const override;
  const B200LinearSolver& linearSolver() const { return *solver_; }

 private:
  mutable std::shared_ptr<B200LinearSolver> solver_;
};

/// The same seam on GaussNewtonOptimizer (GaussNewtonOptimizer.cpp:54).
class B200SolveGaussNewtonOptimizer : public gtsam::GaussNewtonOptimizer {
 public:
  B200SolveGaussNewtonOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                                const gtsam::GaussNewtonParams& params = gtsam::GaussNewtonParams());
  gtsam::VectorValues solve(const gtsam::GaussianFactorGraph& gfg, const gtsam::NonlinearOptimizerParams& params) const override;

 private:
  mutable std::shared_ptr<B200LinearSolver> solver_;
};

/// GaussianFactorGraph::optimize-level entry for the nonlinear graph at `values`:
/// one undamped (lambda = 0) or damped linearize + multifrontal solve on the device.
gtsam::VectorValues solveOnDevice(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& values,
                                  const gtsam::Ordering& ordering, double lambda = 0.0);

}  // namespace gtsam_b200
