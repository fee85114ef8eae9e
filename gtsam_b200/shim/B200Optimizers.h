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
 * BetweenFactor<Pose3>, PriorFactor<Pose3|Point3|PinholeCamera<Cal3Bundler>>,
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

#include <memory>

namespace gtsam_b200 {

struct DeviceState;  // packed problem + C-ABI handles

class B200LevenbergMarquardtOptimizer : public gtsam::LevenbergMarquardtOptimizer {
 public:
  B200LevenbergMarquardtOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                                  const gtsam::LevenbergMarquardtParams& params = gtsam::LevenbergMarquardtParams());
  B200LevenbergMarquardtOptimizer(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& initialValues,
                                  const gtsam::Ordering& ordering,
                                  const gtsam::LevenbergMarquardtParams& params = gtsam::LevenbergMarquardtParams());
  ~B200LevenbergMarquardtOptimizer() override;

  /// One LM iteration on the device.  Returns nullptr: the linear graph stays in HBM
  /// (call linearize() to materialise it on the host).
  gtsam::GaussianFactorGraph::shared_ptr iterate() override;
  gtsam::GaussianFactorGraph::shared_ptr linearize() const override;
  /// kernels launched so far (evidence that the device path ran)
  long long launchCount() const;

 private:
  void init();
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
  /// Marginals::marginalCovariance, gtsam/nonlinear/Marginals.cpp:118-126
  gtsam::Matrix marginalCovariance(gtsam::Key variable) const;
  /// Marginals::marginalInformation, gtsam/nonlinear/Marginals.cpp:128-154
  gtsam::Matrix marginalInformation(gtsam::Key variable) const;
  /// JointMarginal::fullMatrix() of Marginals::jointMarginalCovariance: blocks in sorted-key order
  /// (gtsam::JointMarginal's constructor is private to gtsam::Marginals, hence the plain matrix)
  gtsam::Matrix jointMarginalCovariance(const gtsam::KeyVector& variables) const;

 private:
  std::shared_ptr<DeviceState> dev_;
};

/// GaussianFactorGraph::optimize-level entry for the nonlinear graph at `values`:
/// one undamped (lambda = 0) or damped linearize + multifrontal solve on the device.
gtsam::VectorValues solveOnDevice(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& values,
                                  const gtsam::Ordering& ordering, double lambda = 0.0);

}  // namespace gtsam_b200
