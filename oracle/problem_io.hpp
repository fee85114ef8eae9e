/*
 * problem_io.hpp — TEST INFRASTRUCTURE ONLY: reads a gtsam_b200.problem.Problem.save()
 * file and rebuilds it as real GTSAM objects (NonlinearFactorGraph, Values, Ordering).
 * Shared by oracle/ref_harness.cpp and tests/shim_parity.cpp.
 * Variable id i <-> gtsam::Key i (plain integers), so Key order == id order.
 */
#pragma once
#include <gtsam/geometry/Cal3Bundler.h>
#include <gtsam/geometry/Cal3_S2.h>
#include <gtsam/geometry/PinholeCamera.h>
#include <gtsam/geometry/Point3.h>
#include <gtsam/geometry/Pose2.h>
#include <gtsam/geometry/Pose3.h>
#include <gtsam/inference/Ordering.h>
#include <gtsam/linear/GaussianBayesTree.h>
#include <gtsam/linear/GaussianEliminationTree.h>
#include <gtsam/linear/GaussianJunctionTree.h>
#include <gtsam/linear/JacobianFactor.h>
#include <gtsam/nonlinear/GaussNewtonOptimizer.h>
#include <gtsam/nonlinear/LevenbergMarquardtOptimizer.h>
#include <gtsam/nonlinear/NonlinearFactorGraph.h>
#include <gtsam/nonlinear/PriorFactor.h>
#include <gtsam/nonlinear/Values.h>
#include <gtsam/slam/BetweenFactor.h>
#include <gtsam/slam/GeneralSFMFactor.h>
#include <gtsam/slam/ProjectionFactor.h>
#include <gtsam/sfm/SfmData.h>
#include <gtsam/inference/Symbol.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <vector>

using namespace gtsam;
typedef PinholeCamera<Cal3Bundler> BCam;

static const int VAR_STORAGE[4] = {12, 3, 17, 3};
static const int F_ARITY[8] = {2, 1, 1, 2, 2, 1, 2, 1};
static const int F_MEAS[8] = {12, 12, 3, 2, 2, 17, 3, 3};
static const int F_DIM[8] = {6, 6, 3, 2, 2, 9, 3, 3};

struct Group {
  int32_t type, noise_kind, per_factor, has_cal;
  int64_t count, gi0;
  std::vector<int64_t> keys;
  std::vector<double> meas, noise;
  std::vector<int32_t> cal_index;
  std::vector<double> body;  // body_P_sensor (12) when has_cal & 2
  std::vector<int64_t> gidx; // explicit graph positions when has_cal & 4
  int robust_kind = 0;       // (has_cal >> 8) & 0xff
  double robust_param = 0;
};
struct Prob {
  int64_t nvars;
  std::vector<int32_t> var_type;
  std::vector<double> values;
  std::vector<int64_t> ordering;
  std::vector<double> cal;
  std::vector<Group> groups;
};

template <class T>
static void rd(std::ifstream& f, T* p, size_t n) { f.read((char*)p, (std::streamsize)(n * sizeof(T))); }

static Prob load(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
  char magic[8];
  rd(f, magic, 8);
  if (memcmp(magic, "B200PRB1", 8)) { fprintf(stderr, "bad magic\n"); exit(2); }
  Prob p;
  rd(f, &p.nvars, 1);
  p.var_type.resize(p.nvars);
  rd(f, p.var_type.data(), p.nvars);
  int64_t nval;
  rd(f, &nval, 1);
  p.values.resize(nval);
  rd(f, p.values.data(), nval);
  p.ordering.resize(p.nvars);
  rd(f, p.ordering.data(), p.nvars);
  int64_t ncal;
  rd(f, &ncal, 1);
  p.cal.resize(ncal * 5);
  rd(f, p.cal.data(), ncal * 5);
  int64_t ng;
  rd(f, &ng, 1);
  p.groups.resize(ng);
  for (auto& g : p.groups) {
    rd(f, &g.type, 1); rd(f, &g.noise_kind, 1); rd(f, &g.per_factor, 1); rd(f, &g.has_cal, 1);
    rd(f, &g.count, 1); rd(f, &g.gi0, 1);
    g.keys.resize(g.count * F_ARITY[g.type]);
    rd(f, g.keys.data(), g.keys.size());
    g.meas.resize(g.count * F_MEAS[g.type]);
    rd(f, g.meas.data(), g.meas.size());
    int64_t nn;
    rd(f, &nn, 1);
    g.noise.resize(nn);
    rd(f, g.noise.data(), nn);
    if (g.has_cal & 1) { g.cal_index.resize(g.count); rd(f, g.cal_index.data(), g.count); }
    if (g.has_cal & 2) { g.body.resize(12); rd(f, g.body.data(), 12); }
    g.robust_kind = (g.has_cal >> 8) & 0xff;
    if (g.robust_kind) rd(f, &g.robust_param, 1);
    if (g.has_cal & 4) { g.gidx.resize(g.count); rd(f, g.gidx.data(), g.count); }
  }
  return p;
}

static Pose3 mkpose(const double* x) {
  Matrix3 R;
  R << x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], x[8];
  return Pose3(Rot3(R), Point3(x[9], x[10], x[11]));
}
static BCam mkcam(const double* x) { return BCam(mkpose(x), Cal3Bundler(x[12], x[13], x[14], x[15], x[16])); }
static void putpose(const Pose3& p, double* x) {
  Matrix3 R = p.rotation().matrix();
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) x[3 * i + j] = R(i, j);
  x[9] = p.x(); x[10] = p.y(); x[11] = p.z();
}

static SharedNoiseModel mknoise_base(const Group& g, int64_t i);
static SharedNoiseModel mknoise(const Group& g, int64_t i) {
  SharedNoiseModel base = mknoise_base(g, i);
  switch (g.robust_kind) {
    case 1: return noiseModel::Robust::Create(noiseModel::mEstimator::Huber::Create(g.robust_param), base);
    case 2: return noiseModel::Robust::Create(noiseModel::mEstimator::Cauchy::Create(g.robust_param), base);
    case 3: return noiseModel::Robust::Create(noiseModel::mEstimator::Tukey::Create(g.robust_param), base);
    case 4: return noiseModel::Robust::Create(noiseModel::mEstimator::Fair::Create(g.robust_param), base);
    default: return base;
  }
}
static SharedNoiseModel mknoise_base(const Group& g, int64_t i) {
  const int d = F_DIM[g.type];
  const int pay = g.noise_kind == 0 ? 0 : g.noise_kind == 1 ? 1 : g.noise_kind == 2 ? d : d * d;
  const double* nz = g.noise.data() + (g.per_factor ? i * pay : 0);
  switch (g.noise_kind) {
    case 0: return noiseModel::Unit::Create(d);
    case 1: return noiseModel::Isotropic::Sigma(d, nz[0]);
    case 2: { Vector s(d); for (int k = 0; k < d; k++) s(k) = nz[k]; return noiseModel::Diagonal::Sigmas(s); }
    default: {
      Matrix R(d, d);
      for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) R(r, c) = nz[r * d + c];
      return noiseModel::Gaussian::SqrtInformation(R);
    }
  }
}

struct Built {
  NonlinearFactorGraph graph;
  Values values;
  Ordering ordering;
  std::vector<int64_t> val_off;
};

static Built build(const Prob& p) {
  Built b;
  b.val_off.assign(p.nvars + 1, 0);
  for (int64_t v = 0; v < p.nvars; v++) b.val_off[v + 1] = b.val_off[v] + VAR_STORAGE[p.var_type[v]];
  for (int64_t v = 0; v < p.nvars; v++) {
    const double* x = p.values.data() + b.val_off[v];
    switch (p.var_type[v]) {
      case 0: b.values.insert(Key(v), mkpose(x)); break;
      case 1: b.values.insert(Key(v), Point3(x[0], x[1], x[2])); break;
      case 2: b.values.insert(Key(v), mkcam(x)); break;
      case 3: b.values.insert(Key(v), Pose2(x[0], x[1], x[2])); break;
    }
  }
  for (int64_t j = 0; j < p.nvars; j++) b.ordering.push_back(Key(p.ordering[j]));
  int64_t total = 0;
  for (auto& g : p.groups) total += g.count;
  std::vector<NonlinearFactor::shared_ptr> fs(total);
  std::vector<std::shared_ptr<Cal3_S2>> Ks;
  for (size_t c = 0; c < p.cal.size() / 5; c++) {
    const double* k = p.cal.data() + 5 * c;
    Ks.push_back(std::make_shared<Cal3_S2>(k[0], k[1], k[2], k[3], k[4]));
  }
  for (auto& g : p.groups) {
    SharedNoiseModel shared = g.per_factor ? SharedNoiseModel() : mknoise(g, 0);
    for (int64_t i = 0; i < g.count; i++) {
      SharedNoiseModel nm = g.per_factor ? mknoise(g, i) : shared;
      const int64_t* k = g.keys.data() + i * F_ARITY[g.type];
      const double* z = g.meas.data() + i * F_MEAS[g.type];
      NonlinearFactor::shared_ptr f;
      switch (g.type) {
        case 0: f = std::make_shared<BetweenFactor<Pose3>>(k[0], k[1], mkpose(z), nm); break;
        case 1: f = std::make_shared<PriorFactor<Pose3>>(k[0], mkpose(z), nm); break;
        case 2: f = std::make_shared<PriorFactor<Point3>>(k[0], Point3(z[0], z[1], z[2]), nm); break;
        case 3:
          if (g.has_cal & 2)
            f = std::make_shared<GenericProjectionFactor<Pose3, Point3, Cal3_S2>>(
                Point2(z[0], z[1]), nm, k[0], k[1], Ks[(g.has_cal & 1) ? g.cal_index[i] : 0], mkpose(g.body.data()));
          else
            f = std::make_shared<GenericProjectionFactor<Pose3, Point3, Cal3_S2>>(
                Point2(z[0], z[1]), nm, k[0], k[1], Ks[(g.has_cal & 1) ? g.cal_index[i] : 0]);
          break;
        case 4: f = std::make_shared<GeneralSFMFactor<BCam, Point3>>(Point2(z[0], z[1]), nm, k[0], k[1]); break;
        case 5: f = std::make_shared<PriorFactor<BCam>>(k[0], mkcam(z), nm); break;
        case 6: f = std::make_shared<BetweenFactor<Pose2>>(k[0], k[1], Pose2(z[0], z[1], z[2]), nm); break;
        case 7: f = std::make_shared<PriorFactor<Pose2>>(k[0], Pose2(z[0], z[1], z[2]), nm); break;
      }
      fs[(g.has_cal & 4) ? g.gidx[i] : g.gi0 + i] = f;
    }
  }
  for (auto& f : fs) b.graph.push_back(f);
  return b;
}

