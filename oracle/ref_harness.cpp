/*
 * ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference (oracle/_ref/libgtsam_ref.so, built by
 * oracle/Makefile from the sources under /root/reference) through its own
 * public API on problems written by gtsam_b200.problem.Problem.save().
 * Used to (1) pin the C oracle and generate tests/golden/ fixtures
 * (tests/golden/make_golden.py), (2) produce COLAMD / METIS orderings (inputs
 * at the C-ABI boundary), (3) time the reference's CPU path for bench.py's
 * `--impl reference` arm and `cpu_baseline`.
 *
 * Variable id i <-> gtsam::Key i (plain integers), so Key order == id order.
 */
#include "problem_io.hpp"
#include "linear_io.hpp"
#include <gtsam/nonlinear/DoglegOptimizer.h>
#include <gtsam/nonlinear/Marginals.h>
#include <gtsam/nonlinear/GncOptimizer.h>
#include <gtsam/slam/dataset.h>
#include <gtsam/geometry/Pose2.h>
#include <gtsam/sfm/SfmData.h>
#include <gtsam/inference/Symbol.h>

/* ---- named-array output container ---------------------------------------- */
struct Out {
  std::ofstream f;
  explicit Out(const std::string& path) : f(path, std::ios::binary) { f.write("B200OUT1", 8); }
  void put(const std::string& name, const std::vector<double>& a) { rec(name, 'd', a.data(), a.size(), 8); }
  void put(const std::string& name, const std::vector<int64_t>& a) { rec(name, 'q', a.data(), a.size(), 8); }
  void put(const std::string& name, double x) { put(name, std::vector<double>{x}); }
  void rec(const std::string& name, char t, const void* d, size_t n, size_t sz) {
    int32_t nl = (int32_t)name.size();
    int64_t cnt = (int64_t)n;
    f.write((char*)&nl, 4); f.write(name.data(), nl); f.write(&t, 1); f.write((char*)&cnt, 8);
    f.write((const char*)d, (std::streamsize)(n * sz));
  }
};

static std::vector<double> pack_values(const Prob& p, const Built& b, const Values& v) {
  std::vector<double> out(b.val_off[p.nvars]);
  for (int64_t i = 0; i < p.nvars; i++) {
    double* x = out.data() + b.val_off[i];
    switch (p.var_type[i]) {
      case 0: putpose(v.at<Pose3>(i), x); break;
      case 1: { Point3 q = v.at<Point3>(i); x[0] = q.x(); x[1] = q.y(); x[2] = q.z(); break; }
      case 2: {
        const BCam& c = v.at<BCam>(i);
        putpose(c.pose(), x);
        x[12] = c.calibration().fx(); x[13] = c.calibration().k1(); x[14] = c.calibration().k2();
        x[15] = c.calibration().px(); x[16] = c.calibration().py();
        break;
      }
      case 3: { const Pose2& q = v.at<Pose2>(i); x[0] = q.x(); x[1] = q.y(); x[2] = q.theta(); break; }
    }
  }
  return out;
}

static std::vector<double> pack_vv(const Prob& p, const VectorValues& vv) {
  std::vector<double> out;
  for (int64_t i = 0; i < p.nvars; i++) {
    const Vector& x = vv.at(Key(i));
    for (int k = 0; k < x.size(); k++) out.push_back(x(k));
  }
  return out;
}

/* whitened [A1 A2 b] of every factor of a group, factor-major, col-major blocks */
static void dump_jacobians(const Prob& p, const GaussianFactorGraph& lin, Out& out) {
  for (size_t gi = 0; gi < p.groups.size(); gi++) {
    const Group& g = p.groups[gi];
    std::vector<double> J;
    for (int64_t i = 0; i < g.count; i++) {
      const int64_t gpos = (g.has_cal & 4) ? g.gidx[i] : g.gi0 + i;
      auto jf = std::dynamic_pointer_cast<JacobianFactor>(lin[gpos]);
      if (!jf) { fprintf(stderr, "factor %ld is not a JacobianFactor\n", (long)gpos); exit(3); }
      /* apply any remaining (non-unit) model so the dump is the whitened system */
      Matrix Ab = jf->augmentedJacobian();  // whitened [A b]
      /* columns are in the factor's key order, which for our factors is (key1,key2) */
      for (int c = 0; c < Ab.cols(); c++) for (int r = 0; r < Ab.rows(); r++) J.push_back(Ab(r, c));
    }
    out.put("J" + std::to_string(gi), J);
  }
}

static void dump_tree(const Prob& p, const GaussianFactorGraph& gfg, const Ordering& ordering, Out& out) {
  auto bt = gfg.eliminateMultifrontal(ordering, EliminatePreferCholesky);
  std::vector<int64_t> fptr{0}, fvars, sptr{0}, svars, head;
  std::vector<double> conds;  // [R S d] col-major per clique, in the order dumped
  std::vector<int64_t> cptr{0};
  // traverse
  std::vector<GaussianBayesTree::sharedClique> stack(bt->roots().begin(), bt->roots().end());
  while (!stack.empty()) {
    auto c = stack.back();
    stack.pop_back();
    auto cond = c->conditional();
    for (auto it = cond->beginFrontals(); it != cond->endFrontals(); ++it) fvars.push_back((int64_t)*it);
    for (auto it = cond->beginParents(); it != cond->endParents(); ++it) svars.push_back((int64_t)*it);
    fptr.push_back((int64_t)fvars.size());
    sptr.push_back((int64_t)svars.size());
    Matrix Ab = cond->augmentedJacobian();  // [R S d], unit model
    for (int cc = 0; cc < Ab.cols(); cc++) for (int r = 0; r < Ab.rows(); r++) conds.push_back(Ab(r, cc));
    cptr.push_back((int64_t)conds.size());
    for (auto& ch : c->children) stack.push_back(ch);
  }
  out.put("clique_frontal_ptr", fptr);
  out.put("clique_frontal_vars", fvars);
  out.put("clique_separator_ptr", sptr);
  out.put("clique_separator_vars", svars);
  out.put("clique_cond_ptr", cptr);
  out.put("clique_cond", conds);
}

static int cmd_dump(const std::string& in, const std::string& outp, double lambda, bool diag) {
  Prob p = load(in);
  Built b = build(p);
  Out out(outp);
  out.put("error", b.graph.error(b.values));
  auto lin = b.graph.linearize(b.values);
  dump_jacobians(p, *lin, out);
  out.put("hessian_diagonal", pack_vv(p, lin->hessianDiagonal()));
  LevenbergMarquardtParams params;
  params.ordering = b.ordering;
  params.lambdaInitial = lambda > 0 ? lambda : 1e-5;
  params.diagonalDamping = diag;
  LevenbergMarquardtOptimizer lm(b.graph, b.values, params);
  GaussianFactorGraph sys = *lin;
  if (lambda > 0) {
    VectorValues sqrtHD;
    if (diag) {
      sqrtHD = lin->hessianDiagonal();
      for (auto& kv : sqrtHD) kv.second = kv.second.cwiseMax(params.minDiagonal).cwiseMin(params.maxDiagonal).cwiseSqrt();
    }
    sys = lm.buildDampedSystem(*lin, sqrtHD);
  }
  int status = 0;
  VectorValues delta;
  try {
    delta = sys.optimize(b.ordering, EliminatePreferCholesky);
  } catch (const IndeterminantLinearSystemException& e) {
    status = 1;
    out.put("fail_var", std::vector<int64_t>{(int64_t)e.nearbyVariable()});
  }
  out.put("status", std::vector<int64_t>{status});
  if (!status) {
    out.put("delta", pack_vv(p, delta));
    out.put("linear_error_zero", lin->error(VectorValues::Zero(delta)));
    out.put("linear_error_delta", lin->error(delta));
    Values nv = b.values.retract(delta);
    out.put("new_values", pack_values(p, b, nv));
    out.put("new_error", b.graph.error(nv));
    dump_tree(p, sys, b.ordering, out);
  }
  return 0;
}

static LevenbergMarquardtParams lm_params(const Built& b, bool ceres, int maxit) {
  LevenbergMarquardtParams params = ceres ? LevenbergMarquardtParams::CeresDefaults() : LevenbergMarquardtParams::LegacyDefaults();
  params.ordering = b.ordering;
  params.maxIterations = maxit;
  return params;
}

static int cmd_lm(const std::string& in, const std::string& outp, int maxit, bool ceres) {
  Prob p = load(in);
  Built b = build(p);
  Out out(outp);
  LevenbergMarquardtOptimizer lm(b.graph, b.values, lm_params(b, ceres, maxit));
  std::vector<double> errs{lm.error()}, lams{lm.lambda()};
  std::vector<int64_t> inner{0};
  /* NonlinearOptimizer::defaultOptimize loop, with a trace */
  double currentError, newError = lm.error();
  const auto& prm = lm.params();
  if (!(newError <= prm.errorTol)) {
    do {
      currentError = newError;
      lm.iterate();
      newError = lm.error();
      errs.push_back(newError);
      lams.push_back(lm.lambda());
      inner.push_back(lm.getInnerIterations());
    } while ((int)lm.iterations() < maxit &&
             !checkConvergence(prm.relativeErrorTol, prm.absoluteErrorTol, prm.errorTol, currentError, newError) &&
             std::isfinite(currentError));
  }
  out.put("lm_errors", errs);
  out.put("lm_lambdas", lams);
  out.put("lm_inner", inner);
  out.put("lm_iterations", std::vector<int64_t>{(int64_t)lm.iterations()});
  out.put("final_values", pack_values(p, b, lm.values()));
  return 0;
}

static int cmd_gn(const std::string& in, const std::string& outp, int iters) {
  Prob p = load(in);
  Built b = build(p);
  Out out(outp);
  GaussNewtonParams params;
  params.ordering = b.ordering;
  GaussNewtonOptimizer gn(b.graph, b.values, params);
  std::vector<double> errs{gn.error()};
  for (int i = 0; i < iters; i++) { gn.iterate(); errs.push_back(gn.error()); }
  out.put("gn_errors", errs);
  out.put("final_values", pack_values(p, b, gn.values()));
  return 0;
}

static int cmd_marginals(const std::string& in, const std::string& outp) {
  Prob p = load(in);
  Built b = build(p);
  Out out(outp);
  Marginals marginals(b.graph, b.values, b.ordering, Marginals::CHOLESKY);
  std::vector<double> cov;
  for (int64_t v = 0; v < p.nvars; v++) {
    const Matrix S = marginals.marginalCovariance(Key(v));
    for (int j = 0; j < S.cols(); j++)
      for (int i = 0; i < S.rows(); i++) cov.push_back(S(i, j));   // column-major per variable, variables in id order
  }
  out.put("marg_cov", cov);
  return 0;
}

static int cmd_gnc(const std::string& in, const std::string& outp, int loss) {
  Prob p = load(in);
  Built b = build(p);
  Out out(outp);
  LevenbergMarquardtParams lmp;
  lmp.ordering = b.ordering;
  GncParams<LevenbergMarquardtParams> gp(lmp);
  gp.lossType = loss ? GncLossType::TLS : GncLossType::GM;
  GncOptimizer<GncParams<LevenbergMarquardtParams>> gnc(b.graph, b.values, gp);
  const Values result = gnc.optimize();
  const Vector w = gnc.getWeights(), th = gnc.getInlierCostThresholds();
  out.put("gnc_weights", std::vector<double>(w.data(), w.data() + w.size()));
  out.put("gnc_barcsq", std::vector<double>(th.data(), th.data() + th.size()));
  out.put("final_values", pack_values(p, b, result));
  out.put("final_error", std::vector<double>{b.graph.error(result)});
  return 0;
}

static int cmd_jointmarg(const std::string& in, const std::string& outp, int argc, char** argv) {
  Prob p = load(in);
  Built b = build(p);
  Out out(outp);
  Marginals marginals(b.graph, b.values, b.ordering, Marginals::CHOLESKY);
  KeyVector keys;
  for (int i = 4; i < argc; i++) keys.push_back(Key(atoll(argv[i])));
  const JointMarginal joint = marginals.jointMarginalCovariance(keys);
  const Matrix F = joint.fullMatrix();   // blocks in sorted-key order (Marginals.cpp:176-188)
  std::vector<double> cov;
  for (int j = 0; j < F.cols(); j++)
    for (int i = 0; i < F.rows(); i++) cov.push_back(F(i, j));
  out.put("joint_cov", cov);
  return 0;
}

static int cmd_dogleg(const std::string& in, const std::string& outp, int iters, double delta0) {
  Prob p = load(in);
  Built b = build(p);
  Out out(outp);
  DoglegParams params;
  params.ordering = b.ordering;
  params.deltaInitial = delta0;
  DoglegOptimizer dl(b.graph, b.values, params);
  std::vector<double> errs{dl.error()}, deltas{dl.getDelta()};
  for (int i = 0; i < iters; i++) { dl.iterate(); errs.push_back(dl.error()); deltas.push_back(dl.getDelta()); }
  out.put("dl_errors", errs);
  out.put("dl_deltas", deltas);
  out.put("final_values", pack_values(p, b, dl.values()));
  return 0;
}

static double now() {
  return std::chrono::duration<double>(std::chrono::high_resolution_clock::now().time_since_epoch()).count();
}

/* time LM iterate() of the stock reference: each step = fresh optimizer on the
 * initial values (construction untimed) + one iterate() (timed) */
static int cmd_time(const std::string& in, int steps, int warmup, bool ceres) {
  Prob p = load(in);
  Built b = build(p);
  std::vector<double> ts, tl, tsv;
  double err_after = 0;
  int inner = 0;
  for (int s = 0; s < warmup + steps; s++) {
    LevenbergMarquardtOptimizer lm(b.graph, b.values, lm_params(b, ceres, 100));
    double t0 = now();
    lm.iterate();
    double t1 = now();
    if (s >= warmup) ts.push_back(t1 - t0);
    err_after = lm.error();
    inner = lm.getInnerIterations();
  }
  { /* split: linearize and one damped solve */
    double t0 = now();
    auto lin = b.graph.linearize(b.values);
    double t1 = now();
    LevenbergMarquardtOptimizer lm(b.graph, b.values, lm_params(b, ceres, 100));
    auto sys = lm.buildDampedSystem(*lin, VectorValues());
    double t2 = now();
    auto d = sys.optimize(b.ordering, EliminatePreferCholesky);
    double t3 = now();
    tl.push_back(t1 - t0);
    tsv.push_back(t3 - t2);
  }
  double sum = 0, mx = 0;
  for (double t : ts) { sum += t; mx = std::max(mx, t); }
  printf("{\"steps\": %d, \"warmup\": %d, \"total_s\": %.6f, \"mean_s\": %.6f, \"max_s\": %.6f, "
         "\"linearize_s\": %.6f, \"solve_s\": %.6f, \"error_after\": %.12g, \"inner_iterations\": %d, "
         "\"nfactors\": %zu, \"nvars\": %ld}\n",
         steps, warmup, sum, sum / std::max<size_t>(1, ts.size()), mx, tl[0], tsv[0], err_after, inner,
         b.graph.size(), (long)p.nvars);
  return 0;
}

static int cmd_order(const std::string& in, const std::string& kind, const std::string& outp) {
  Prob p = load(in);
  Built b = build(p);
  Ordering ord;
  if (kind == "colamd") ord = Ordering::Colamd(b.graph);
  else if (kind == "metis") ord = Ordering::Metis(b.graph);
  else if (kind == "natural") ord = Ordering::Natural(b.graph);
  else { fprintf(stderr, "unknown ordering %s\n", kind.c_str()); return 2; }
  std::vector<int64_t> o;
  for (Key k : ord) o.push_back((int64_t)k);
  Out out(outp);
  out.put("ordering", o);
  return 0;
}

template <class T>
static void wr(std::ofstream& f, const T* p, size_t n) { f.write((const char*)p, (std::streamsize)(n * sizeof(T))); }

static void save(const Prob& p, const std::string& path) {
  std::ofstream f(path, std::ios::binary);
  f.write("B200PRB1", 8);
  wr(f, &p.nvars, 1);
  wr(f, p.var_type.data(), p.var_type.size());
  int64_t nval = (int64_t)p.values.size();
  wr(f, &nval, 1);
  wr(f, p.values.data(), p.values.size());
  wr(f, p.ordering.data(), p.ordering.size());
  int64_t ncal = (int64_t)p.cal.size() / 5;
  wr(f, &ncal, 1);
  wr(f, p.cal.data(), p.cal.size());
  int64_t ng = (int64_t)p.groups.size();
  wr(f, &ng, 1);
  for (auto& g : p.groups) {
    wr(f, &g.type, 1); wr(f, &g.noise_kind, 1); wr(f, &g.per_factor, 1); wr(f, &g.has_cal, 1);
    wr(f, &g.count, 1); wr(f, &g.gi0, 1);
    wr(f, g.keys.data(), g.keys.size());
    wr(f, g.meas.data(), g.meas.size());
    int64_t nn = (int64_t)g.noise.size();
    wr(f, &nn, 1);
    wr(f, g.noise.data(), g.noise.size());
    if (g.has_cal & 1) wr(f, g.cal_index.data(), g.cal_index.size());
    if (g.has_cal & 2) wr(f, g.body.data(), g.body.size());
    if ((g.has_cal >> 8) & 0xff) wr(f, &g.robust_param, 1);
    if (g.has_cal & 4) wr(f, g.gidx.data(), g.gidx.size());
  }
}

/* Convert a BAL file through the reference's own loader (gtsam/sfm/SfmData.cpp:189-246)
 * into a problem file.  mode 0: tests/testGeneralSFMFactorB.cpp:44-63 (unit noise,
 * default COLAMD ordering); mode 1: examples/SFMExample_bal.cpp:36-78 (isotropic
 * noise + priors on camera 0 and point 0, COLAMD). */
static int cmd_balfile(const std::string& path, const std::string& outp, int mode) {
  SfmData db = SfmData::FromBalFile(path);
  const int64_t nc = (int64_t)db.numberCameras(), np = (int64_t)db.numberTracks();
  Prob p;
  p.nvars = nc + np;
  p.var_type.assign(nc, 2);
  p.var_type.insert(p.var_type.end(), np, 1);
  NonlinearFactorGraph graph;
  auto camkey = [&](size_t i) { return mode == 0 ? Key(i) : Key(Symbol('c', i)); };
  auto ptkey = [&](size_t j) { return Key(Symbol('p', j)); };
  for (auto& cam : db.cameras) {
    double x[17];
    putpose(cam.pose(), x);
    x[12] = cam.calibration().fx(); x[13] = cam.calibration().k1(); x[14] = cam.calibration().k2();
    x[15] = cam.calibration().px(); x[16] = cam.calibration().py();
    p.values.insert(p.values.end(), x, x + 17);
  }
  for (auto& t : db.tracks) { p.values.push_back(t.p.x()); p.values.push_back(t.p.y()); p.values.push_back(t.p.z()); }
  Group g;
  g.type = 4; g.noise_kind = mode == 0 ? 0 : 1; g.per_factor = 0; g.has_cal = 0; g.gi0 = 0;
  if (mode == 1) g.noise.push_back(1.0);
  auto noise = mode == 0 ? SharedNoiseModel(noiseModel::Unit::Create(2)) : SharedNoiseModel(noiseModel::Isotropic::Sigma(2, 1.0));
  for (size_t j = 0; j < db.numberTracks(); j++)
    for (const SfmMeasurement& m : db.tracks[j].measurements) {
      graph.emplace_shared<GeneralSFMFactor<BCam, Point3>>(m.second, noise, camkey(m.first), ptkey(j));
      g.keys.push_back((int64_t)m.first); g.keys.push_back(nc + (int64_t)j);
      g.meas.push_back(m.second.x()); g.meas.push_back(m.second.y());
    }
  g.count = (int64_t)g.meas.size() / 2;
  p.groups.push_back(g);
  if (mode == 1) {
    graph.addPrior(camkey(0), db.cameras[0], noiseModel::Isotropic::Sigma(9, 0.1));
    graph.addPrior(ptkey(0), db.tracks[0].p, noiseModel::Isotropic::Sigma(3, 0.1));
    Group gc; gc.type = 5; gc.noise_kind = 1; gc.per_factor = 0; gc.has_cal = 0; gc.count = 1; gc.gi0 = g.count;
    gc.keys = {0}; gc.meas.assign(p.values.begin(), p.values.begin() + 17); gc.noise = {0.1};
    Group gp; gp.type = 2; gp.noise_kind = 1; gp.per_factor = 0; gp.has_cal = 0; gp.count = 1; gp.gi0 = g.count + 1;
    gp.keys = {nc}; gp.meas = {db.tracks[0].p.x(), db.tracks[0].p.y(), db.tracks[0].p.z()}; gp.noise = {0.1};
    p.groups.push_back(gc); p.groups.push_back(gp);
  }
  Ordering ord = Ordering::Colamd(graph);
  for (Key k : ord) {
    Symbol s(k);
    int64_t id = (mode == 0 && k < (Key)nc) ? (int64_t)k : (s.chr() == 'c' ? (int64_t)s.index() : nc + (int64_t)s.index());
    p.ordering.push_back(id);
  }
  save(p, outp);
  return 0;
}

/* Convert a g2o 3D pose graph through the reference's own loader (gtsam/slam/dataset.cpp:922-944)
 * into a problem file, with the prior of examples/Pose3SLAMExample_g2o.cpp:42-49 and the
 * COLAMD ordering GaussNewtonOptimizer would compute. */
static int cmd_g2ofile(const std::string& path, const std::string& outp) {
  auto [graph, initial] = readG2o(path, true);
  auto priorModel = noiseModel::Diagonal::Variances((Vector(6) << 1e-6, 1e-6, 1e-6, 1e-4, 1e-4, 1e-4).finished());
  const Key firstKey = initial->keys().front();
  graph->addPrior(firstKey, Pose3(), priorModel);
  Prob p;
  std::map<Key, int64_t> id;
  for (const auto& kv : *initial) { id[kv.key] = (int64_t)id.size(); }
  p.nvars = (int64_t)id.size();
  p.var_type.assign(p.nvars, 0);
  for (const auto& kv : *initial) {
    double x[12];
    putpose(initial->at<Pose3>(kv.key), x);
    p.values.insert(p.values.end(), x, x + 12);
  }
  Group gb; gb.type = 0; gb.noise_kind = 3; gb.per_factor = 1; gb.has_cal = 0; gb.gi0 = 0; gb.count = 0;
  for (const auto& f : *graph) {
    auto b = std::dynamic_pointer_cast<BetweenFactor<Pose3>>(f);
    if (!b) continue;
    gb.keys.push_back(id[b->key1()]); gb.keys.push_back(id[b->key2()]);
    double x[12];
    putpose(b->measured(), x);
    gb.meas.insert(gb.meas.end(), x, x + 12);
    const Matrix R = std::dynamic_pointer_cast<noiseModel::Gaussian>(b->noiseModel())->R();
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) gb.noise.push_back(R(r, c));
    gb.count++;
  }
  Group gp; gp.type = 1; gp.noise_kind = 2; gp.per_factor = 0; gp.has_cal = 0; gp.count = 1; gp.gi0 = gb.count;
  gp.keys = {id[firstKey]};
  { double x[12]; putpose(Pose3(), x); gp.meas.assign(x, x + 12); }
  for (double v : {1e-6, 1e-6, 1e-6, 1e-4, 1e-4, 1e-4}) gp.noise.push_back(std::sqrt(v));
  p.groups = {gb, gp};
  Ordering ord = Ordering::Colamd(*graph);
  for (Key k : ord) p.ordering.push_back(id[k]);
  save(p, outp);
  return 0;
}

/* BASELINE.json configs[0]: Pose2SLAMExample_g2o on a small g2o file, CPU only (plumbing proof that
 * the reference built by oracle/Makefile runs its own example path).  Mirrors
 * examples/Pose2SLAMExample_g2o.cpp:46-84 with GaussNewton (as shipped) and with
 * LevenbergMarquardt (as configs[0] names it); prints one JSON line. */
static int cmd_pose2(const std::string& path) {
  auto [graph, initial] = readG2o(path, false);
  auto priorModel = noiseModel::Diagonal::Variances(Vector3(1e-6, 1e-6, 1e-8));
  graph->addPrior(0, Pose2(), priorModel);
  const double e0 = graph->error(*initial);
  GaussNewtonParams gp;
  GaussNewtonOptimizer gn(*graph, *initial, gp);
  Values rg = gn.optimize();
  LevenbergMarquardtOptimizer lm(*graph, *initial);
  Values rl = lm.optimize();
  printf("{\"file\": \"%s\", \"variables\": %zu, \"factors\": %zu, \"initial_error\": %.12g, "
         "\"gn_final_error\": %.12g, \"gn_iterations\": %d, \"lm_final_error\": %.12g, \"lm_iterations\": %d}\n",
         path.substr(path.find_last_of('/') + 1).c_str(), initial->size(), graph->size(), e0, graph->error(rg),
         (int)gn.iterations(), graph->error(rl), (int)lm.iterations());
  return 0;
}

/* ---- GaussianFactorGraph level: the reference's own optimize() on a graph of JacobianFactors ----
 * `linsolve in.lin out [lambda]`: GaussianFactorGraph::optimize(ordering, EliminatePreferCholesky)
 * (gtsam/linear/GaussianFactorGraph.cpp:316-319); with lambda > 0 the graph is first extended by the
 * damping priors buildDampedSystem appends (LevenbergMarquardtState.h:125-156, non-diagonal case:
 * A = I, b = 0, sigma = 1/sqrt(lambda) per variable).  Dumps delta, hessianDiagonal, the two linear
 * errors of the UNDAMPED graph, the Bayes tree (cliques + conditionals) and the marginal covariance of
 * every variable (inverse of the marginal information, GaussianBayesTree::marginalFactor). */
static int cmd_linsolve(const std::string& in, const std::string& outp, double lambda) {
  linio::LinProb lp = linio::load(in);
  GaussianFactorGraph gfg = linio::build_graph(lp);
  Ordering ordering = linio::build_ordering(lp);
  Out out(outp);
  Prob dummy;
  dummy.nvars = lp.nvars;
  auto pack = [&](const VectorValues& vv) {
    std::vector<double> o;
    for (int64_t i = 0; i < lp.nvars; i++) { const Vector& x = vv.at(Key(i)); for (int k = 0; k < x.size(); k++) o.push_back(x(k)); }
    return o;
  };
  out.put("hessian_diagonal", pack(gfg.hessianDiagonal()));
  GaussianFactorGraph sys = gfg;
  if (lambda > 0) {
    for (int64_t v = 0; v < lp.nvars; v++) {
      const int d = lp.var_dim[v];
      sys.push_back(std::make_shared<JacobianFactor>(Key(v), Matrix::Identity(d, d), Vector::Zero(d),
                                                     noiseModel::Isotropic::Sigma(d, 1.0 / std::sqrt(lambda))));
    }
  }
  int status = 0;
  VectorValues delta;
  try {
    delta = sys.optimize(ordering, EliminatePreferCholesky);
  } catch (const IndeterminantLinearSystemException& e) {
    status = 1;
    out.put("fail_var", std::vector<int64_t>{(int64_t)e.nearbyVariable()});
  }
  out.put("status", std::vector<int64_t>{status});
  if (!status) {
    out.put("delta", pack(delta));
    out.put("linear_error_zero", gfg.error(VectorValues::Zero(delta)));
    out.put("linear_error_delta", gfg.error(delta));
    dump_tree(dummy, sys, ordering, out);
    if (lambda == 0) {
      auto bt = gfg.eliminateMultifrontal(ordering, EliminatePreferCholesky);
      std::vector<double> cov;
      for (int64_t v = 0; v < lp.nvars; v++) {
        Matrix info = bt->marginalFactor(Key(v), EliminatePreferCholesky)->information();
        Matrix c = info.inverse();
        for (int cc = 0; cc < c.cols(); cc++) for (int r = 0; r < c.rows(); r++) cov.push_back(c(r, cc));
      }
      out.put("marginal_covariances", cov);
    }
  }
  return 0;
}

/* `linearize2d file.g2o out.lin`: BASELINE.json configs[0]'s graph (Pose2 g2o + the example's prior,
 * examples/Pose2SLAMExample_g2o.cpp:46-64) linearized by the reference at the file's initial estimate,
 * written as a linear problem with the reference's COLAMD ordering. */
static int cmd_linearize2d(const std::string& path, const std::string& outp) {
  auto [graph, initial] = readG2o(path, false);
  auto priorModel = noiseModel::Diagonal::Variances(Vector3(1e-6, 1e-6, 1e-8));
  graph->addPrior(0, Pose2(), priorModel);
  auto lin = graph->linearize(*initial);
  Ordering ordering = Ordering::Colamd(*lin);
  linio::LinProb lp;
  if (!linio::from_graph(*lin, ordering, &lp)) { fprintf(stderr, "graph holds a non-Jacobian or constrained factor\n"); return 3; }
  linio::save(lp, outp);
  printf("{\"variables\": %ld, \"factors\": %ld, \"groups\": %zu, \"error\": %.12g}\n", (long)lp.nvars, (long)lp.nfactors(),
         lp.groups.size(), graph->error(*initial));
  return 0;
}

/* known-answer vectors for the geometry primitives, incl. near-0 / near-pi */
static int cmd_kat(const std::string& outp) {
  std::mt19937 rng(123);
  std::normal_distribution<double> N(0, 1);
  std::vector<double> xi_in, T_out, log_out, ad_out, inv_out, comp_out, so3_w, so3_R, so3_log;
  std::vector<Vector6> xis;
  for (int i = 0; i < 40; i++) { Vector6 x; for (int k = 0; k < 6; k++) x(k) = N(rng); xis.push_back(x); }
  for (double s : {1e-12, 1e-9, 1e-7, 1e-4}) { Vector6 x; for (int k = 0; k < 6; k++) x(k) = N(rng); x.head<3>() *= s; xis.push_back(x); }
  for (double eps : {0.0, 1e-9, 1e-6, 1e-4, 1e-2}) {  // rotations near pi about varied axes
    for (int a = 0; a < 4; a++) {
      Vector3 ax(N(rng), N(rng), N(rng));
      if (a < 3) { ax = Vector3::Zero(); ax(a) = 1; }
      ax.normalize();
      Vector6 x; x << ax * (M_PI - eps), N(rng), N(rng), N(rng);
      xis.push_back(x);
    }
  }
  Pose3 prev;
  for (auto& x : xis) {
    Pose3 T = Pose3::Expmap(x);
    double buf[12];
    for (int k = 0; k < 6; k++) xi_in.push_back(x(k));
    putpose(T, buf); T_out.insert(T_out.end(), buf, buf + 12);
    Vector6 l = Pose3::Logmap(T);
    for (int k = 0; k < 6; k++) log_out.push_back(l(k));
    Matrix6 Ad = T.AdjointMap();
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) ad_out.push_back(Ad(r, c));
    putpose(T.inverse(), buf); inv_out.insert(inv_out.end(), buf, buf + 12);
    putpose(prev * T, buf); comp_out.insert(comp_out.end(), buf, buf + 12);
    prev = T;
    Vector3 w = x.head<3>();
    Rot3 R = Rot3::Expmap(w);
    Vector3 lw = Rot3::Logmap(R);
    for (int k = 0; k < 3; k++) { so3_w.push_back(w(k)); so3_log.push_back(lw(k)); }
    Matrix3 Rm = R.matrix();
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) so3_R.push_back(Rm(r, c));
  }
  Out out(outp);
  out.put("xi", xi_in); out.put("expmap", T_out); out.put("logmap", log_out); out.put("adjoint", ad_out);
  out.put("inverse", inv_out); out.put("compose_prev", comp_out);
  out.put("so3_w", so3_w); out.put("so3_R", so3_R); out.put("so3_log", so3_log);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: ref_harness dump|lm|gn|time|order|kat ...\n"); return 2; }
  std::string cmd = argv[1];
  if (cmd == "dump" && argc >= 4) return cmd_dump(argv[2], argv[3], argc > 4 ? atof(argv[4]) : 0.0, argc > 5 && atoi(argv[5]));
  if (cmd == "lm" && argc >= 4) return cmd_lm(argv[2], argv[3], argc > 4 ? atoi(argv[4]) : 100, argc > 5 && atoi(argv[5]));
  if (cmd == "gnc" && argc >= 5) return cmd_gnc(argv[2], argv[3], atoi(argv[4]));
  if (cmd == "jointmarg" && argc >= 5) return cmd_jointmarg(argv[2], argv[3], argc, argv);
  if (cmd == "marginals" && argc >= 4) return cmd_marginals(argv[2], argv[3]);
  if (cmd == "dogleg" && argc >= 4) return cmd_dogleg(argv[2], argv[3], argc > 4 ? atoi(argv[4]) : 5, argc > 5 ? atof(argv[5]) : 1.0);
  if (cmd == "gn" && argc >= 4) return cmd_gn(argv[2], argv[3], argc > 4 ? atoi(argv[4]) : 3);
  if (cmd == "time" && argc >= 3) return cmd_time(argv[2], argc > 3 ? atoi(argv[3]) : 3, argc > 4 ? atoi(argv[4]) : 1, argc > 5 && atoi(argv[5]));
  if (cmd == "order" && argc >= 5) return cmd_order(argv[2], argv[3], argv[4]);
  if (cmd == "kat" && argc >= 3) return cmd_kat(argv[2]);
  if (cmd == "pose2" && argc >= 3) return cmd_pose2(argv[2]);
  if (cmd == "linsolve" && argc >= 4) return cmd_linsolve(argv[2], argv[3], argc > 4 ? atof(argv[4]) : 0.0);
  if (cmd == "linearize2d" && argc >= 4) return cmd_linearize2d(argv[2], argv[3]);
  if (cmd == "g2ofile" && argc >= 4) return cmd_g2ofile(argv[2], argv[3]);
  if (cmd == "balfile" && argc >= 4) return cmd_balfile(argv[2], argv[3], argc > 4 ? atoi(argv[4]) : 0);
  fprintf(stderr, "bad arguments\n");
  return 2;
}
