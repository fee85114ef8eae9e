// B200Optimizers.cpp — see B200Optimizers.h.  Host glue only: packs the GTSAM
// objects once, then every numeric step is a C-ABI call into libgtsam_b200.so.
#include "B200Optimizers.h"

#include <gtsam/geometry/Cal3Bundler.h>
#include <gtsam/geometry/Cal3_S2.h>
#include <gtsam/geometry/PinholeCamera.h>
#include <gtsam/geometry/Pose2.h>
#include <gtsam/geometry/Pose3.h>
#include <gtsam/linear/HessianFactor.h>
#include <gtsam/linear/JacobianFactor.h>
#include <gtsam/linear/linearExceptions.h>
#include <gtsam/nonlinear/PriorFactor.h>
#include <gtsam/nonlinear/internal/LevenbergMarquardtState.h>
#include <gtsam/nonlinear/internal/NonlinearOptimizerState.h>
#include <gtsam/slam/BetweenFactor.h>
#include <gtsam/slam/GeneralSFMFactor.h>
#include <gtsam/slam/ProjectionFactor.h>

#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gtsam_b200.h"

using namespace gtsam;

namespace gtsam_b200 {

typedef PinholeCamera<Cal3Bundler> BCam;
typedef GenericProjectionFactor<Pose3, Point3, Cal3_S2> ProjFactor;
typedef GeneralSFMFactor<BCam, Point3> SfmFactor;

static void check(int rc, const char* what) {
  if (rc == B200_OK) return;
  throw std::runtime_error(std::string("gtsam_b200: ") + what + ": " + b200_last_error_string());
}

static void putPose(const Pose3& p, std::vector<double>& out) {
  const Matrix3 R = p.rotation().matrix();
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out.push_back(R(i, j));
  out.push_back(p.x()); out.push_back(p.y()); out.push_back(p.z());
}
static void putCam(const BCam& c, std::vector<double>& out) {
  putPose(c.pose(), out);
  const Cal3Bundler& k = c.calibration();
  out.push_back(k.fx()); out.push_back(k.k1()); out.push_back(k.k2()); out.push_back(k.px()); out.push_back(k.py());
}
static Pose3 getPose(const double* x) {
  Matrix3 R;
  R << x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], x[8];
  return Pose3(Rot3(R), Point3(x[9], x[10], x[11]));
}

struct GroupBuf {
  int type, noise_kind;
  int robust_kind = 0;
  double robust_param = 0;
  std::vector<int64_t> keys;
  std::vector<double> meas, noise;
  std::vector<int32_t> cal;
  std::vector<double> body;   // body_P_sensor shared by the group (empty: none)
  std::vector<int64_t> pos;   // graph position of every factor (factors of one kind need not be consecutive)
  int64_t count = 0;
};

struct DeviceState {
  std::vector<Key> id2key;
  std::map<Key, int64_t> key2id;
  std::vector<int32_t> var_type;
  std::vector<int64_t> val_off, dof_off;
  std::vector<GroupBuf> groups;
  std::vector<double> cal;
  b200_ctx* ctx = nullptr;
  b200_problem* prob = nullptr;
  b200_lm* lm = nullptr;
  B200Communicator comm;      // world == 1: single GPU

  ~DeviceState() {
    if (lm) b200_lm_destroy(lm);
    if (prob) b200_problem_destroy(prob);
    if (ctx) b200_ctx_destroy(ctx);
  }

  std::vector<double> packValues(const Values& values) {
    std::vector<double> out;
    for (Key k : id2key) {
      const Value& v = values.at(k);
      if (auto p = dynamic_cast<const GenericValue<Pose3>*>(&v)) putPose(p->value(), out);
      else if (auto q = dynamic_cast<const GenericValue<Point3>*>(&v)) { out.push_back(q->value().x()); out.push_back(q->value().y()); out.push_back(q->value().z()); }
      else if (auto c = dynamic_cast<const GenericValue<BCam>*>(&v)) putCam(c->value(), out);
      else if (auto p2 = dynamic_cast<const GenericValue<Pose2>*>(&v)) { out.push_back(p2->value().x()); out.push_back(p2->value().y()); out.push_back(p2->value().theta()); }
      else throw std::invalid_argument("gtsam_b200: unsupported Value type for key " + DefaultKeyFormatter(k));
    }
    return out;
  }

  Values unpackValues(const std::vector<double>& x) const {
    Values out;
    for (size_t i = 0; i < id2key.size(); i++) {
      const double* v = x.data() + val_off[i];
      switch (var_type[i]) {
        case B200_VAR_POSE3: out.insert(id2key[i], getPose(v)); break;
        case B200_VAR_POINT3: out.insert(id2key[i], Point3(v[0], v[1], v[2])); break;
        case B200_VAR_CAM_BUNDLER: out.insert(id2key[i], BCam(getPose(v), Cal3Bundler(v[12], v[13], v[14], v[15], v[16]))); break;
        case B200_VAR_POSE2: out.insert(id2key[i], Pose2(v[0], v[1], v[2])); break;
      }
    }
    return out;
  }

  // noise model -> (kind, payload); Constrained / Robust are rejected like an unsupported factor
  // noiseModel::Robust = m-estimator around a base model (gtsam/linear/NoiseModel.h: Robust)
  static SharedNoiseModel unwrapRobust(const SharedNoiseModel& nm, int* kind, double* param) {
    *kind = B200_ROBUST_NONE; *param = 0;
    auto rb = std::dynamic_pointer_cast<noiseModel::Robust>(nm);
    if (!rb) return nm;
    auto m = rb->robust();
    if (auto h = std::dynamic_pointer_cast<noiseModel::mEstimator::Huber>(m)) { *kind = B200_ROBUST_HUBER; *param = h->modelParameter(); }
    else if (auto c = std::dynamic_pointer_cast<noiseModel::mEstimator::Cauchy>(m)) { *kind = B200_ROBUST_CAUCHY; *param = c->modelParameter(); }
    else if (auto t = std::dynamic_pointer_cast<noiseModel::mEstimator::Tukey>(m)) { *kind = B200_ROBUST_TUKEY; *param = t->modelParameter(); }
    else if (auto f = std::dynamic_pointer_cast<noiseModel::mEstimator::Fair>(m)) { *kind = B200_ROBUST_FAIR; *param = f->modelParameter(); }
    else throw std::invalid_argument("gtsam_b200: unsupported m-estimator (supported: Huber, Cauchy, Tukey, Fair)");
    return rb->noise();
  }

  static int noiseOf(const SharedNoiseModel& nm, int d, std::vector<double>& payload) {
    payload.clear();
    if (!nm || nm->isUnit()) return B200_NOISE_UNIT;
    if (nm->isConstrained()) throw std::invalid_argument("gtsam_b200: Constrained noise models are out of scope (need QR)");
    if (auto iso = std::dynamic_pointer_cast<noiseModel::Isotropic>(nm)) { payload.push_back(iso->sigma()); return B200_NOISE_ISOTROPIC; }
    if (auto dg = std::dynamic_pointer_cast<noiseModel::Diagonal>(nm)) {
      for (int i = 0; i < d; i++) payload.push_back(dg->sigma(i));
      return B200_NOISE_DIAGONAL;
    }
    if (auto g = std::dynamic_pointer_cast<noiseModel::Gaussian>(nm)) {
      const Matrix R = g->R();
      for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) payload.push_back(R(r, c));
      return B200_NOISE_GAUSSIAN;
    }
    throw std::invalid_argument("gtsam_b200: unsupported noise model");
  }

  void pack(const NonlinearFactorGraph& graph, const Values& values, const Ordering& ordering) {
    // ids in ascending Key order == iteration order of Values (gtsam/nonlinear/Values.h:74-79)
    for (const auto& kv : values) {
      key2id[kv.key] = (int64_t)id2key.size();
      id2key.push_back(kv.key);
    }
    val_off.assign(1, 0); dof_off.assign(1, 0);
    for (Key k : id2key) {
      const Value& v = values.at(k);
      int t;
      if (dynamic_cast<const GenericValue<Pose3>*>(&v)) t = B200_VAR_POSE3;
      else if (dynamic_cast<const GenericValue<Point3>*>(&v)) t = B200_VAR_POINT3;
      else if (dynamic_cast<const GenericValue<BCam>*>(&v)) t = B200_VAR_CAM_BUNDLER;
      else if (dynamic_cast<const GenericValue<Pose2>*>(&v)) t = B200_VAR_POSE2;
      else throw std::invalid_argument("gtsam_b200: unsupported Value type for key " + DefaultKeyFormatter(k));
      var_type.push_back(t);
      val_off.push_back(val_off.back() + b200_var_storage(t));
      dof_off.push_back(dof_off.back() + b200_var_dim(t));
    }
    std::map<const Cal3_S2*, int32_t> calIds;
    std::vector<double> pay;
    int64_t pos = 0;
    for (const auto& f : graph) {
      if (!f) throw std::invalid_argument("gtsam_b200: null factors are not supported");
      int type, d;
      std::vector<int64_t> keys;
      std::vector<double> meas;
      int32_t cal_id = 0;
      std::vector<double> body;
      SharedNoiseModel nm;
      auto id = [&](Key k) {
        auto it = key2id.find(k);
        if (it == key2id.end()) throw ValuesKeyDoesNotExist("gtsam_b200 pack", k);
        return it->second;
      };
      if (auto b = dynamic_cast<const BetweenFactor<Pose3>*>(f.get())) {
        type = B200_FACTOR_BETWEEN_POSE3; keys = {id(b->key1()), id(b->key2())}; putPose(b->measured(), meas); nm = b->noiseModel();
      } else if (auto b2 = dynamic_cast<const BetweenFactor<Pose2>*>(f.get())) {
        type = B200_FACTOR_BETWEEN_POSE2; keys = {id(b2->key1()), id(b2->key2())};
        meas = {b2->measured().x(), b2->measured().y(), b2->measured().theta()}; nm = b2->noiseModel();
      } else if (auto q2 = dynamic_cast<const PriorFactor<Pose2>*>(f.get())) {
        type = B200_FACTOR_PRIOR_POSE2; keys = {id(q2->key())};
        meas = {q2->prior().x(), q2->prior().y(), q2->prior().theta()}; nm = q2->noiseModel();
      } else if (auto p3 = dynamic_cast<const PriorFactor<Pose3>*>(f.get())) {
        type = B200_FACTOR_PRIOR_POSE3; keys = {id(p3->key())}; putPose(p3->prior(), meas); nm = p3->noiseModel();
      } else if (auto pp = dynamic_cast<const PriorFactor<Point3>*>(f.get())) {
        type = B200_FACTOR_PRIOR_POINT3; keys = {id(pp->key())};
        meas = {pp->prior().x(), pp->prior().y(), pp->prior().z()}; nm = pp->noiseModel();
      } else if (auto pc = dynamic_cast<const PriorFactor<BCam>*>(f.get())) {
        type = B200_FACTOR_PRIOR_CAM_BUNDLER; keys = {id(pc->key())}; putCam(pc->prior(), meas); nm = pc->noiseModel();
      } else if (auto pj = dynamic_cast<const ProjFactor*>(f.get())) {
        if (pj->body_P_sensor()) putPose(*pj->body_P_sensor(), body);
        type = B200_FACTOR_PROJECTION_CAL3S2; keys = {id(pj->key1()), id(pj->key2())};
        meas = {pj->measured().x(), pj->measured().y()}; nm = pj->noiseModel();
        const Cal3_S2* K = pj->calibration().get();
        auto it = calIds.find(K);
        if (it == calIds.end()) {
          it = calIds.emplace(K, (int32_t)calIds.size()).first;
          cal.insert(cal.end(), {K->fx(), K->fy(), K->skew(), K->px(), K->py()});
        }
        cal_id = it->second;
      } else if (auto sf = dynamic_cast<const SfmFactor*>(f.get())) {
        type = B200_FACTOR_SFM_BUNDLER; keys = {id(sf->key1()), id(sf->key2())};
        meas = {sf->measured().x(), sf->measured().y()}; nm = sf->noiseModel();
      } else {
        throw std::invalid_argument("gtsam_b200: unsupported factor type at graph position " + std::to_string(pos) +
                                    " (no CPU fallback; supported: Between<Pose3|Pose2>, Prior<Pose3|Pose2|Point3|SfmCamera>, "
                                    "GenericProjectionFactor<Pose3,Point3,Cal3_S2>, GeneralSFMFactor<SfmCamera,Point3>)");
      }
      d = b200_factor_dim(type);
      int rkind; double rparam;
      const SharedNoiseModel base = unwrapRobust(nm, &rkind, &rparam);
      const int kind = noiseOf(base, d, pay);
      // bucket by (type, noise kind, robust loss, body_P_sensor): a graph that interleaves factor
      // kinds still becomes a handful of homogeneous tables, each with explicit graph positions
      size_t gidx = groups.size();
      for (size_t q = 0; q < groups.size(); q++)
        if (groups[q].type == type && groups[q].noise_kind == kind && groups[q].body == body &&
            groups[q].robust_kind == rkind && groups[q].robust_param == rparam) { gidx = q; break; }
      if (gidx == groups.size()) {
        GroupBuf g; g.type = type; g.noise_kind = kind; g.body = body; g.robust_kind = rkind; g.robust_param = rparam;
        groups.push_back(g);
      }
      GroupBuf& g = groups[gidx];
      g.pos.push_back(pos);
      g.keys.insert(g.keys.end(), keys.begin(), keys.end());
      g.meas.insert(g.meas.end(), meas.begin(), meas.end());
      g.noise.insert(g.noise.end(), pay.begin(), pay.end());
      g.cal.push_back(cal_id);
      g.count++;
      pos++;
    }
    // ---- C-ABI description ----
    std::vector<int64_t> ord;
    for (Key k : ordering) {
      auto it = key2id.find(k);
      if (it == key2id.end()) throw std::invalid_argument("gtsam_b200: ordering contains a key that is not in Values");
      ord.push_back(it->second);
    }
    if (ord.size() != id2key.size()) throw std::invalid_argument("gtsam_b200: ordering must cover every variable");
    std::vector<b200_factor_group> cg(groups.size());
    for (size_t i = 0; i < groups.size(); i++) {
      cg[i].type = groups[i].type; cg[i].noise_kind = groups[i].noise_kind;
      cg[i].noise_per_factor = groups[i].noise_kind != B200_NOISE_UNIT && groups[i].count > 1;
      cg[i].robust_kind = groups[i].robust_kind; cg[i].robust_param = groups[i].robust_param; cg[i].count = groups[i].count; cg[i].graph_index0 = -1; cg[i].graph_index = groups[i].pos.data();
      cg[i].keys = groups[i].keys.data(); cg[i].meas = groups[i].meas.data(); cg[i].noise = groups[i].noise.data();
      cg[i].cal_index = groups[i].type == B200_FACTOR_PROJECTION_CAL3S2 ? groups[i].cal.data() : nullptr;
      cg[i].body_P_sensor = groups[i].body.empty() ? nullptr : groups[i].body.data();
    }
    const std::vector<double> packed = packValues(values);
    b200_problem_desc desc;
    desc.nvars = (int64_t)id2key.size(); desc.var_type = var_type.data(); desc.values = packed.data();
    desc.ordering = ord.data(); desc.ncal = (int64_t)cal.size() / 5; desc.cal = cal.data();
    desc.ngroups = (int64_t)cg.size(); desc.groups = cg.data();
    const char* devEnv = std::getenv("B200_DEVICE");
    check(b200_ctx_create(comm.world > 1 ? comm.device : (devEnv ? std::atoi(devEnv) : 0), &ctx), "b200_ctx_create");
    if (comm.world > 1) check(b200_ctx_comm_init(ctx, comm.uniqueId.data(), comm.rank, comm.world), "b200_ctx_comm_init");
    check(b200_problem_create(ctx, &desc, &prob), "b200_problem_create");
  }

  Values currentValues() const {
    std::vector<double> x((size_t)b200_values_size(prob));
    // sharded: every rank owns a part of the new estimate; b200_get_values_all gathers the whole of it on every rank
    check(comm.world > 1 ? b200_get_values_all(prob, x.data()) : b200_get_values(prob, x.data()), "b200_get_values");
    return unpackValues(x);
  }

  VectorValues currentDelta() const {
    std::vector<double> dl((size_t)b200_delta_size(prob));
    check(b200_get_delta(prob, dl.data()), "b200_get_delta");
    VectorValues out;
    for (size_t i = 0; i < id2key.size(); i++)
      out.insert(id2key[i], Eigen::Map<const Vector>(dl.data() + dof_off[i], dof_off[i + 1] - dof_off[i]));
    return out;
  }
};

static b200_lm_params toC(const LevenbergMarquardtParams& p) {
  b200_lm_params c;
  c.max_iterations = (int)p.maxIterations; c.relative_error_tol = p.relativeErrorTol;
  c.absolute_error_tol = p.absoluteErrorTol; c.error_tol = p.errorTol; c.lambda_initial = p.lambdaInitial;
  c.lambda_factor = p.lambdaFactor; c.lambda_upper_bound = p.lambdaUpperBound; c.lambda_lower_bound = p.lambdaLowerBound;
  c.min_model_fidelity = p.minModelFidelity; c.diagonal_damping = p.diagonalDamping;
  c.use_fixed_lambda_factor = p.useFixedLambdaFactor; c.min_diagonal = p.minDiagonal; c.max_diagonal = p.maxDiagonal;
  return c;
}

// ---- Levenberg-Marquardt ---------------------------------------------------------
B200LevenbergMarquardtOptimizer::B200LevenbergMarquardtOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues,
                                                                 const LevenbergMarquardtParams& params)
    : LevenbergMarquardtOptimizer(graph, initialValues, params) { init(); }
B200LevenbergMarquardtOptimizer::B200LevenbergMarquardtOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues,
                                                                 const Ordering& ordering, const LevenbergMarquardtParams& params)
    : LevenbergMarquardtOptimizer(graph, initialValues, ordering, params) { init(); }
B200LevenbergMarquardtOptimizer::B200LevenbergMarquardtOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues,
                                                                 const Ordering& ordering, const LevenbergMarquardtParams& params,
                                                                 const B200Communicator& comm)
    : LevenbergMarquardtOptimizer(graph, initialValues, ordering, params) { init(&comm); }
B200LevenbergMarquardtOptimizer::~B200LevenbergMarquardtOptimizer() {}

std::array<char, 128> B200Communicator::newUniqueId() {
  std::array<char, 128> id{};
  check(b200_nccl_unique_id(id.data()), "b200_nccl_unique_id");
  return id;
}

void B200LevenbergMarquardtOptimizer::init(const B200Communicator* comm) {
  dev_ = std::make_shared<DeviceState>();
  if (comm) dev_->comm = *comm;
  dev_->pack(graph_, state_->values, *params_.ordering);  // ordering is always set by the base ctor
  const b200_lm_params c = toC(params_);
  check(b200_lm_create(dev_->prob, &c, &dev_->lm), "b200_lm_create");
}

GaussianFactorGraph::shared_ptr B200LevenbergMarquardtOptimizer::iterate() {
  check(b200_lm_iterate(dev_->lm), "b200_lm_iterate");
  b200_lm_state s;
  b200_lm_get_state(dev_->lm, &s);
  typedef internal::LevenbergMarquardtState State;
  state_.reset(new State(dev_->currentValues(), s.error, s.lambda, s.current_factor, (unsigned)s.iterations,
                         (unsigned)s.total_inner_iterations));
  return GaussianFactorGraph::shared_ptr();
}

GaussianFactorGraph::shared_ptr B200LevenbergMarquardtOptimizer::linearize() const {
  if (dev_->comm.world > 1) throw std::invalid_argument("gtsam_b200: linearize() of a sharded optimizer would hold this rank's factors only");
  check(b200_linearize(dev_->prob), "b200_linearize");
  size_t total = 0;
  for (auto& g : dev_->groups) total += (size_t)g.count;
  std::vector<GaussianFactor::shared_ptr> out(total);
  for (size_t gi = 0; gi < dev_->groups.size(); gi++) {
    const GroupBuf& g = dev_->groups[gi];
    const int d = b200_factor_dim(g.type), ar = b200_factor_arity(g.type);
    int ncols = 1;
    std::vector<int> dims;
    for (int a = 0; a < ar; a++) { dims.push_back(b200_var_dim(dev_->var_type[g.keys[a]])); ncols += dims.back(); }
    std::vector<double> J((size_t)g.count * d * ncols);
    check(b200_get_jacobians(dev_->prob, (int64_t)gi, J.data()), "b200_get_jacobians");
    for (int64_t i = 0; i < g.count; i++) {
      Eigen::Map<const Matrix> Ab(J.data() + (size_t)i * d * ncols, d, ncols);
      const Vector b = Ab.col(ncols - 1);
      const Key k1 = dev_->id2key[g.keys[i * ar]];
      if (ar == 1) out[g.pos[i]] = std::make_shared<JacobianFactor>(k1, Matrix(Ab.leftCols(dims[0])), b);
      else out[g.pos[i]] = std::make_shared<JacobianFactor>(k1, Matrix(Ab.leftCols(dims[0])), dev_->id2key[g.keys[i * ar + 1]],
                                                             Matrix(Ab.middleCols(dims[0], dims[1])), b);
    }
  }
  auto gfg = std::make_shared<GaussianFactorGraph>();
  for (auto& f : out) gfg->push_back(f);
  return gfg;
}

long long B200LevenbergMarquardtOptimizer::launchCount() const { return b200_launch_count(dev_->ctx); }
void B200LevenbergMarquardtOptimizer::setJacobianFp32(bool on) { check(b200_set_jacobian_precision(dev_->prob, on ? 1 : 0), "b200_set_jacobian_precision"); }

// ---- Gauss-Newton ------------------------------------------------------------------
B200GaussNewtonOptimizer::B200GaussNewtonOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues,
                                                   const GaussNewtonParams& params)
    : GaussNewtonOptimizer(graph, initialValues, params) { init(); }
B200GaussNewtonOptimizer::B200GaussNewtonOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues,
                                                   const Ordering& ordering)
    : GaussNewtonOptimizer(graph, initialValues, ordering) { init(); }
B200GaussNewtonOptimizer::~B200GaussNewtonOptimizer() {}

void B200GaussNewtonOptimizer::init() {
  dev_ = std::make_shared<DeviceState>();
  dev_->pack(graph_, state_->values, *params_.ordering);
}

GaussianFactorGraph::shared_ptr B200GaussNewtonOptimizer::iterate() {
  double e = 0;
  const int rc = b200_gn_iterate(dev_->prob, &e);
  if (rc == B200_INDETERMINATE) throw IndeterminantLinearSystemException(0);
  check(rc, "b200_gn_iterate");
  state_.reset(new internal::NonlinearOptimizerState(dev_->currentValues(), e, state_->iterations + 1));
  return GaussianFactorGraph::shared_ptr();
}

// ---- Dogleg -------------------------------------------------------------------------------
static DoglegParams doglegWithOrdering(DoglegParams params, const NonlinearFactorGraph& graph) {
  if (!params.ordering) params.ordering = Ordering::Create(params.orderingType, graph);   // DoglegOptimizer.cpp:126-130
  return params;
}
B200DoglegOptimizer::B200DoglegOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues, const DoglegParams& params)
    : NonlinearOptimizer(graph, std::unique_ptr<internal::NonlinearOptimizerState>(
                                    new internal::NonlinearOptimizerState(initialValues, graph.error(initialValues)))),
      params_(doglegWithOrdering(params, graph)) { init(); }
B200DoglegOptimizer::B200DoglegOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues, const Ordering& ordering)
    : NonlinearOptimizer(graph, std::unique_ptr<internal::NonlinearOptimizerState>(
                                    new internal::NonlinearOptimizerState(initialValues, graph.error(initialValues)))) {
  params_.ordering = ordering;
  init();
}
B200DoglegOptimizer::~B200DoglegOptimizer() {
  if (dl_) b200_dl_destroy((b200_dl*)dl_);
}

void B200DoglegOptimizer::init() {
  dev_ = std::make_shared<DeviceState>();
  dev_->pack(graph_, state_->values, *params_.ordering);
  b200_dl* dl = nullptr;
  check(b200_dl_create(dev_->prob, params_.deltaInitial, &dl), "b200_dl_create");
  dl_ = dl;
}

double B200DoglegOptimizer::getDelta() const {
  double delta = 0;
  check(b200_dl_get_state((const b200_dl*)dl_, nullptr, &delta, nullptr), "b200_dl_get_state");
  return delta;
}

GaussianFactorGraph::shared_ptr B200DoglegOptimizer::iterate() {
  const int rc = b200_dl_iterate((b200_dl*)dl_);
  if (rc == B200_INDETERMINATE) throw IndeterminantLinearSystemException(0);
  check(rc, "b200_dl_iterate");
  double e = 0;
  check(b200_dl_get_state((const b200_dl*)dl_, &e, nullptr, nullptr), "b200_dl_get_state");
  state_.reset(new internal::NonlinearOptimizerState(dev_->currentValues(), e, state_->iterations + 1));
  return GaussianFactorGraph::shared_ptr();
}

VectorValues solveOnDevice(const NonlinearFactorGraph& graph, const Values& values, const Ordering& ordering, double lambda) {
  DeviceState dev;
  dev.pack(graph, values, ordering);
  check(b200_linearize(dev.prob), "b200_linearize");
  double e0, e1;
  int64_t fv = -1;
  const int rc = b200_solve(dev.prob, lambda, 0, 1e-6, 1e32, &e0, &e1, &fv);
  if (rc == B200_INDETERMINATE) throw IndeterminantLinearSystemException(fv >= 0 ? dev.id2key[(size_t)fv] : 0);
  check(rc, "b200_solve");
  return dev.currentDelta();
}

// ---- GaussianFactorGraph level ---------------------------------------------------------------
// JacobianFactors bucketed by shape (rows, has-model, block widths); keys -> dense ids in ascending Key order.
struct LinGroup {
  int rows = 0;            // JacobianFactor rows; N + 1 for a HessianFactor group
  bool hessian = false;    // Ab holds augmented information matrices (HessianFactor::info())
  bool has_model = false;
  std::vector<int32_t> dims;
  std::vector<int64_t> keys, pos;
  std::vector<double> Ab, sigmas;
  int64_t count = 0;
};
struct LinearState {
  std::vector<Key> id2key;
  std::map<Key, int64_t> key2id;
  std::vector<int32_t> var_dim;
  std::vector<int64_t> dof_off;
  std::vector<LinGroup> groups;
  std::vector<std::vector<int>> signature;   // per graph position: (rows, has-model, key ids...)
  std::vector<int64_t> ord;
  std::vector<b200_jacobian_group> cg;
  std::vector<b200_hessian_group> ch;
  std::vector<size_t> jgroups, hgroups;   // indices into `groups` by kind, in the order of the C-ABI arrays
  b200_linear_desc desc;
  b200_ctx* ctx = nullptr;
  b200_problem* prob = nullptr;
  int builds = 0, solves = 0;

  ~LinearState() { reset(); if (ctx) b200_ctx_destroy(ctx); }
  void reset() { if (prob) { b200_problem_destroy(prob); prob = nullptr; } }

  // JacobianFactor (Unit / Diagonal model) or HessianFactor; anything else has no device path
  static void checkFactor(const GaussianFactor::shared_ptr& f, size_t pos) {
    if (!f) throw std::invalid_argument("gtsam_b200: null factor in the GaussianFactorGraph at position " + std::to_string(pos));
    if (auto jf = std::dynamic_pointer_cast<JacobianFactor>(f)) {
      const SharedDiagonal& m = jf->get_model();
      if (m && m->isConstrained()) throw std::invalid_argument("gtsam_b200: Constrained noise models are out of scope (need QR)");
      return;
    }
    if (std::dynamic_pointer_cast<HessianFactor>(f)) return;
    throw std::invalid_argument("gtsam_b200: only JacobianFactors and HessianFactors are supported at the linear level (position " +
                                std::to_string(pos) + " holds another GaussianFactor type); no CPU fallback");
  }
  // shape signature of a factor: (rows or -1 for a HessianFactor, has-model, key ids...)
  std::vector<int> signatureOf(const GaussianFactor::shared_ptr& f, bool* ok) const {
    std::vector<int> sig;
    *ok = true;
    if (auto jf = std::dynamic_pointer_cast<JacobianFactor>(f)) {
      const SharedDiagonal& m = jf->get_model();
      sig = {(int)jf->rows(), (int)(m && !m->isUnit())};
    } else {
      sig = {-1, 0};
    }
    for (auto it = f->begin(); it != f->end(); ++it) {
      auto id = key2id.find(*it);
      if (id == key2id.end() || (int)f->getDim(it) != var_dim[id->second]) { *ok = false; return sig; }
      sig.push_back((int)id->second);
    }
    return sig;
  }

  bool sameStructure(const GaussianFactorGraph& gfg) const {
    if (!prob || gfg.size() != signature.size()) return false;
    for (size_t pos = 0; pos < gfg.size(); pos++) {
      checkFactor(gfg[pos], pos);
      bool ok;
      if (signatureOf(gfg[pos], &ok) != signature[pos] || !ok) return false;
    }
    return true;
  }

  // numbers of every group, in the packed order
  void fillNumbers(const GaussianFactorGraph& gfg) {
    for (auto& g : groups) { g.Ab.clear(); g.sigmas.clear(); }
    for (auto& g : groups)
      for (int64_t pos : g.pos) {
        if (g.hessian) {
          const Matrix info = std::static_pointer_cast<HessianFactor>(gfg[(size_t)pos])->info().selfadjointView();
          g.Ab.insert(g.Ab.end(), info.data(), info.data() + info.size());
          continue;
        }
        auto jf = std::static_pointer_cast<JacobianFactor>(gfg[(size_t)pos]);
        const Matrix Ab = jf->augmentedJacobianUnweighted();   // [A1 .. Ak b], unwhitened; whitening is a device kernel
        g.Ab.insert(g.Ab.end(), Ab.data(), Ab.data() + Ab.size());   // Eigen default is column-major
        if (g.has_model) { const Vector s = jf->get_model()->sigmas(); g.sigmas.insert(g.sigmas.end(), s.data(), s.data() + s.size()); }
      }
  }

  // the C-ABI description of `gfg` (b200_linear_desc over this object's own buffers)
  void pack(const GaussianFactorGraph& gfg, const Ordering& ordering) {
    id2key.clear(); key2id.clear(); var_dim.clear(); dof_off.assign(1, 0); groups.clear(); signature.clear();
    std::map<Key, int> dimOf;
    for (size_t pos = 0; pos < gfg.size(); pos++) {
      checkFactor(gfg[pos], pos);
      const auto& f = gfg[pos];
      for (auto it = f->begin(); it != f->end(); ++it) {
        auto ins = dimOf.emplace(*it, (int)f->getDim(it));
        if (!ins.second && ins.first->second != (int)f->getDim(it))
          throw std::invalid_argument("gtsam_b200: variable " + DefaultKeyFormatter(*it) + " appears with two different dimensions");
      }
    }
    for (auto& kv : dimOf) {   // std::map iterates in ascending Key order
      key2id[kv.first] = (int64_t)id2key.size();
      id2key.push_back(kv.first);
      var_dim.push_back(kv.second);
      dof_off.push_back(dof_off.back() + kv.second);
    }
    std::map<std::vector<int>, size_t> bucket;
    for (size_t pos = 0; pos < gfg.size(); pos++) {
      const auto& f = gfg[pos];
      if (f->size() < 1 || f->size() > B200_JACOBIAN_MAX_ARITY)
        throw std::invalid_argument("gtsam_b200: Gaussian factor with " + std::to_string(f->size()) + " keys (supported: 1.." +
                                    std::to_string(B200_JACOBIAN_MAX_ARITY) + ")");
      bool ok;
      const std::vector<int> sig = signatureOf(f, &ok);
      signature.push_back(sig);
      std::vector<int> shape{sig[0], sig[1]};
      for (auto it = f->begin(); it != f->end(); ++it) shape.push_back((int)f->getDim(it));
      auto found = bucket.find(shape);
      if (found == bucket.end()) {
        LinGroup g;
        g.hessian = sig[0] < 0; g.has_model = sig[1] != 0;
        for (auto it = f->begin(); it != f->end(); ++it) g.dims.push_back((int32_t)f->getDim(it));
        int n1 = 1;
        for (int32_t d : g.dims) n1 += d;
        g.rows = g.hessian ? n1 : sig[0];
        groups.push_back(g);
        found = bucket.emplace(shape, groups.size() - 1).first;
      }
      LinGroup& g = groups[found->second];
      for (auto it = f->begin(); it != f->end(); ++it) g.keys.push_back(key2id.at(*it));
      g.pos.push_back((int64_t)pos);
      g.count++;
    }
    fillNumbers(gfg);
    ord.clear();
    for (Key k : ordering) {
      auto it = key2id.find(k);
      if (it == key2id.end()) throw std::invalid_argument("gtsam_b200: ordering contains a key that is not in the graph");
      ord.push_back(it->second);
    }
    if (ord.size() != id2key.size()) throw std::invalid_argument("gtsam_b200: ordering must cover every variable of the graph");
    cg.clear(); ch.clear(); jgroups.clear(); hgroups.clear();
    for (size_t i = 0; i < groups.size(); i++) {
      const LinGroup& g = groups[i];
      if (g.hessian) {
        b200_hessian_group h;
        h.arity = (int32_t)g.dims.size(); h.dims = g.dims.data(); h.count = g.count; h.graph_index0 = -1;
        h.graph_index = g.pos.data(); h.keys = g.keys.data(); h.info = g.Ab.data();
        ch.push_back(h); hgroups.push_back(i);
      } else {
        b200_jacobian_group c;
        c.rows = g.rows; c.arity = (int32_t)g.dims.size(); c.dims = g.dims.data(); c.count = g.count;
        c.graph_index0 = -1; c.graph_index = g.pos.data(); c.keys = g.keys.data(); c.Ab = g.Ab.data();
        c.sigmas = g.has_model ? g.sigmas.data() : nullptr;
        cg.push_back(c); jgroups.push_back(i);
      }
    }
    desc.nvars = (int64_t)id2key.size(); desc.var_dim = var_dim.data(); desc.ordering = ord.data();
    desc.ngroups = (int64_t)cg.size(); desc.groups = cg.data();
    desc.nhgroups = (int64_t)ch.size(); desc.hgroups = ch.data();
  }

  void build(const GaussianFactorGraph& gfg, const Ordering& ordering) {
    reset();
    pack(gfg, ordering);
    if (!ctx) {
      const char* devEnv = std::getenv("B200_DEVICE");
      check(b200_ctx_create(devEnv ? std::atoi(devEnv) : 0, &ctx), "b200_ctx_create");
    }
    check(b200_linear_create(ctx, &desc, &prob), "b200_linear_create");
    builds++;
  }

  void update(const GaussianFactorGraph& gfg) {
    fillNumbers(gfg);
    for (size_t q = 0; q < jgroups.size(); q++) {
      const LinGroup& g = groups[jgroups[q]];
      check(b200_linear_update(prob, (int64_t)q, g.Ab.data(), g.has_model ? g.sigmas.data() : nullptr), "b200_linear_update");
    }
    for (size_t q = 0; q < hgroups.size(); q++)
      check(b200_linear_update_hessian(prob, (int64_t)q, groups[hgroups[q]].Ab.data()), "b200_linear_update_hessian");
  }

  VectorValues solve() {
    double e0, e1;
    int64_t fv = -1;
    const int rc = b200_solve(prob, 0.0, 0, 0.0, 0.0, &e0, &e1, &fv);
    if (rc == B200_INDETERMINATE) throw IndeterminantLinearSystemException(fv >= 0 ? id2key[(size_t)fv] : 0);
    check(rc, "b200_solve");
    solves++;
    std::vector<double> dl((size_t)b200_delta_size(prob));
    check(b200_get_delta(prob, dl.data()), "b200_get_delta");
    VectorValues out;
    for (size_t i = 0; i < id2key.size(); i++)
      out.insert(id2key[i], Eigen::Map<const Vector>(dl.data() + dof_off[i], dof_off[i + 1] - dof_off[i]));
    return out;
  }

  VectorValues gradientAtZero() {
    std::vector<double> g((size_t)b200_delta_size(prob));
    check(b200_gradient_at_zero(prob, g.data()), "b200_gradient_at_zero");
    VectorValues out;
    for (size_t i = 0; i < id2key.size(); i++)
      out.insert(id2key[i], Eigen::Map<const Vector>(g.data() + dof_off[i], dof_off[i + 1] - dof_off[i]));
    return out;
  }
};

// ---- Marginals ----------------------------------------------------------------------------
B200Marginals::B200Marginals(const NonlinearFactorGraph& graph, const Values& solution, const Ordering& ordering)
    : dev_(std::make_shared<DeviceState>()) {
  dev_->pack(graph, solution, ordering);
}
B200Marginals::B200Marginals(const NonlinearFactorGraph& graph, const Values& solution)
    : B200Marginals(graph, solution, Ordering::Colamd(graph)) {}
B200Marginals::B200Marginals(const GaussianFactorGraph& graph, const Ordering& ordering) : lin_(std::make_shared<LinearState>()) {
  lin_->build(graph, ordering);
}
B200Marginals::B200Marginals(const GaussianFactorGraph& graph) : B200Marginals(graph, Ordering::Colamd(graph)) {}

// (problem handle, id of a key, tangent dimension of an id) of whichever description this object holds
static b200_problem* margProb(const std::shared_ptr<DeviceState>& d, const std::shared_ptr<LinearState>& l) { return d ? d->prob : l->prob; }
static int64_t margId(const std::shared_ptr<DeviceState>& d, const std::shared_ptr<LinearState>& l, Key k, const char* who) {
  const std::map<Key, int64_t>& m = d ? d->key2id : l->key2id;
  const auto it = m.find(k);
  if (it == m.end()) throw ValuesKeyDoesNotExist(who, k);
  return it->second;
}
static int margDim(const std::shared_ptr<DeviceState>& d, const std::shared_ptr<LinearState>& l, int64_t v) {
  return d ? (int)(d->dof_off[v + 1] - d->dof_off[v]) : (int)l->var_dim[(size_t)v];
}

Matrix B200Marginals::marginalCovariance(Key variable) const {
  const int64_t v = margId(dev_, lin_, variable, "B200Marginals::marginalCovariance");
  const int d = margDim(dev_, lin_, v);
  Matrix S(d, d);   // Eigen default: column-major, as the C-ABI writes it
  const int rc = b200_marginal_covariance(margProb(dev_, lin_), v, S.data());
  if (rc == B200_INDETERMINATE) throw IndeterminantLinearSystemException(variable);
  check(rc, "b200_marginal_covariance");
  return S;
}

Matrix B200Marginals::marginalInformation(Key variable) const { return marginalCovariance(variable).inverse(); }

Matrix B200Marginals::jointMarginalCovariance(const KeyVector& variables) const {
  std::vector<int64_t> ids;
  for (Key k : variables) ids.push_back(margId(dev_, lin_, k, "B200Marginals::jointMarginalCovariance"));
  std::sort(ids.begin(), ids.end());   // ids are the ranks of the Keys: sorted ids == sorted keys
  int64_t D = 0;
  for (int64_t v : ids) D += margDim(dev_, lin_, v);
  Matrix S(D, D);
  const int rc = b200_joint_marginal_covariance(margProb(dev_, lin_), ids.data(), (int64_t)ids.size(), S.data());
  if (rc == B200_INDETERMINATE) throw IndeterminantLinearSystemException(variables.front());
  check(rc, "b200_joint_marginal_covariance");
  return S;
}

B200LinearSolver::B200LinearSolver(const Ordering& ordering) : ordering_(ordering), st_(std::make_shared<LinearState>()) {}
B200LinearSolver::~B200LinearSolver() {}
VectorValues B200LinearSolver::optimize(const GaussianFactorGraph& gfg) {
  if (st_->sameStructure(gfg)) st_->update(gfg);
  else st_->build(gfg, ordering_);
  return st_->solve();
}
VectorValues B200LinearSolver::gradientAtZero(const GaussianFactorGraph& gfg) {
  if (st_->sameStructure(gfg)) st_->update(gfg);
  else st_->build(gfg, ordering_);
  return st_->gradientAtZero();
}
int B200LinearSolver::structureBuilds() const { return st_->builds; }
int B200LinearSolver::solves() const { return st_->solves; }
long long B200LinearSolver::launchCount() const { return st_->ctx ? b200_launch_count(st_->ctx) : 0; }

VectorValues optimizeOnDevice(const GaussianFactorGraph& gfg, const Ordering& ordering) {
  B200LinearSolver solver(ordering);
  return solver.optimize(gfg);
}

std::vector<std::pair<KeyVector, KeyVector>> symbolicOnHost(const GaussianFactorGraph& gfg, const Ordering& ordering) {
  LinearState st;
  st.pack(gfg, ordering);
  b200_symbolic* sym = nullptr;
  check(b200_linear_symbolic_create(&st.desc, &sym), "b200_linear_symbolic_create");
  b200_symbolic_info info;
  b200_symbolic_get_info(sym, &info);
  std::vector<int64_t> fp(info.ncliques + 1), sp(info.ncliques + 1), fv(std::max<int64_t>(1, info.frontal_list_len)),
      sv(std::max<int64_t>(1, info.separator_list_len)), par(std::max<int64_t>(1, info.ncliques));
  b200_symbolic_get_cliques(sym, fp.data(), fv.data(), sp.data(), sv.data(), par.data());
  b200_symbolic_destroy(sym);
  std::vector<std::pair<KeyVector, KeyVector>> out;
  for (int64_t c = 0; c < info.ncliques; c++) {
    KeyVector f, s;
    for (int64_t q = fp[c]; q < fp[c + 1]; q++) f.push_back(st.id2key[(size_t)fv[q]]);
    for (int64_t q = sp[c]; q < sp[c + 1]; q++) s.push_back(st.id2key[(size_t)sv[q]]);
    out.emplace_back(f, s);
  }
  return out;
}

GaussianBayesTree::shared_ptr bayesTreeFromTables(const std::vector<KeyVector>& frontals, const std::vector<KeyVector>& separators,
                                                  const std::vector<int64_t>& parent, const std::vector<Matrix>& conditionals,
                                                  const std::map<Key, int>& dims) {
  const size_t nc = frontals.size();
  if (separators.size() != nc || parent.size() != nc || conditionals.size() != nc)
    throw std::invalid_argument("gtsam_b200::bayesTreeFromTables: table sizes differ");
  auto bt = std::make_shared<GaussianBayesTree>();
  std::vector<GaussianBayesTree::sharedClique> cliques(nc);
  for (size_t k = nc; k-- > 0;) {   // top down: parents have larger indices than their children
    KeyVector keys(frontals[k]);
    keys.insert(keys.end(), separators[k].begin(), separators[k].end());
    std::vector<DenseIndex> bd;
    for (Key key : keys) bd.push_back(dims.at(key));
    bd.push_back(1);
    auto cond = std::make_shared<GaussianConditional>(keys, frontals[k].size(), VerticalBlockMatrix(bd, conditionals[k]));
    cliques[k] = std::make_shared<GaussianBayesTreeClique>(cond);
    if (parent[k] >= 0 && (size_t)parent[k] <= k) throw std::invalid_argument("gtsam_b200::bayesTreeFromTables: a parent must follow its children");
    bt->addClique(cliques[k], parent[k] >= 0 ? cliques[(size_t)parent[k]] : GaussianBayesTree::sharedClique());
  }
  return bt;
}

GaussianBayesTree::shared_ptr eliminateMultifrontalOnDevice(const GaussianFactorGraph& gfg, const Ordering& ordering) {
  LinearState st;
  st.build(gfg, ordering);
  st.solve();   // eliminates every clique (and back-substitutes); throws IndeterminantLinearSystemException like the reference
  b200_symbolic_info info;
  check(b200_symbolic_info_get(st.prob, &info), "b200_symbolic_info_get");
  std::vector<int64_t> fp(info.ncliques + 1), sp(info.ncliques + 1), fv(std::max<int64_t>(1, info.frontal_list_len)),
      sv(std::max<int64_t>(1, info.separator_list_len)), par(std::max<int64_t>(1, info.ncliques));
  check(b200_get_cliques(st.prob, fp.data(), fv.data(), sp.data(), sv.data(), par.data()), "b200_get_cliques");
  std::vector<KeyVector> F(info.ncliques), S(info.ncliques);
  std::vector<Matrix> C(info.ncliques);
  std::map<Key, int> dims;
  for (size_t i = 0; i < st.id2key.size(); i++) dims[st.id2key[i]] = st.var_dim[i];
  par.resize(info.ncliques);
  for (int64_t c = 0; c < info.ncliques; c++) {
    int f = 0, s = 0;
    for (int64_t q = fp[c]; q < fp[c + 1]; q++) { F[c].push_back(st.id2key[(size_t)fv[q]]); f += st.var_dim[(size_t)fv[q]]; }
    for (int64_t q = sp[c]; q < sp[c + 1]; q++) { S[c].push_back(st.id2key[(size_t)sv[q]]); s += st.var_dim[(size_t)sv[q]]; }
    C[c].resize(f, f + s + 1);   // Eigen default is column-major, as b200_get_conditional writes it
    check(b200_get_conditional(st.prob, c, C[c].data()), "b200_get_conditional");
  }
  return bayesTreeFromTables(F, S, par, C, dims);
}

static void requireMultifrontalCholesky(const NonlinearOptimizerParams& params) {
  if (params.linearSolverType != NonlinearOptimizerParams::MULTIFRONTAL_CHOLESKY)
    throw std::invalid_argument("gtsam_b200: the device solve() is multifrontal Cholesky "
                                "(NonlinearOptimizerParams::MULTIFRONTAL_CHOLESKY); other linearSolverType values have no device path");
}

B200SolveLevenbergMarquardtOptimizer::B200SolveLevenbergMarquardtOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues,
                                                                           const LevenbergMarquardtParams& params)
    : LevenbergMarquardtOptimizer(graph, initialValues, params) {}
B200SolveLevenbergMarquardtOptimizer::B200SolveLevenbergMarquardtOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues,
                                                                           const Ordering& ordering, const LevenbergMarquardtParams& params)
    : LevenbergMarquardtOptimizer(graph, initialValues, ordering, params) {}
VectorValues B200SolveLevenbergMarquardtOptimizer::solve(const GaussianFactorGraph& gfg, const NonlinearOptimizerParams& params) const {
  requireMultifrontalCholesky(params);
  if (!solver_) solver_ = std::make_shared<B200LinearSolver>(*params.ordering);   // set by the base constructor (ensureHasOrdering)
  return solver_->optimize(gfg);
}

B200SolveGaussNewtonOptimizer::B200SolveGaussNewtonOptimizer(const NonlinearFactorGraph& graph, const Values& initialValues,
                                                             const GaussNewtonParams& params)
    : GaussNewtonOptimizer(graph, initialValues, params) {}
VectorValues B200SolveGaussNewtonOptimizer::solve(const GaussianFactorGraph& gfg, const NonlinearOptimizerParams& params) const {
  requireMultifrontalCholesky(params);
  if (!solver_) solver_ = std::make_shared<B200LinearSolver>(*params.ordering);
  return solver_->optimize(gfg);
}

}  // namespace gtsam_b200
