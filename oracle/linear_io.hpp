/*
 * linear_io.hpp — TEST INFRASTRUCTURE ONLY: the file form of a b200_linear_desc
 * (gtsam_b200.linear.LinearProblem.save()) <-> a real gtsam::GaussianFactorGraph of
 * JacobianFactors + Ordering.  Shared by oracle/ref_harness.cpp and tests/shim_linear.cpp.
 * Variable id i <-> gtsam::Key i.
 *
 * Layout ("B200LIN1"): nvars(q) var_dim(i32 x nvars) ordering(q x nvars) ngroups(q), then per group
 * rows(i32) arity(i32) dims(i32 x arity) count(q) graph_index0(q) flags(i32: 1 sigmas, 4 graph_index,
 * 8 HessianFactor group: rows = sum(dims)+1 and "Ab" holds the augmented information matrices)
 * keys(q x count*arity) Ab(f64 x count*rows*ncols, column-major blocks) [sigmas(f64 x count*rows)]
 * [graph_index(q x count)].
 */
#pragma once
#include <gtsam/inference/Ordering.h>
#include <gtsam/linear/GaussianFactorGraph.h>
#include <gtsam/linear/HessianFactor.h>
#include <gtsam/linear/JacobianFactor.h>
#include <gtsam/linear/NoiseModel.h>
#include <gtsam/linear/VectorValues.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

namespace linio {

struct JGroup {
  int32_t rows = 0, arity = 0, flags = 0;
  std::vector<int32_t> dims;
  int64_t count = 0, gi0 = 0;
  std::vector<int64_t> keys, gidx;
  std::vector<double> Ab, sigmas;
  int ncols() const { int n = 1; for (int d : dims) n += d; return n; }
};
struct LinProb {
  int64_t nvars = 0;
  std::vector<int32_t> var_dim;
  std::vector<int64_t> ordering;
  std::vector<JGroup> groups;
  int64_t nfactors() const { int64_t t = 0; for (auto& g : groups) t += g.count; return t; }
};

template <class T>
static void rdv(std::ifstream& f, T* p, size_t n) { f.read((char*)p, (std::streamsize)(n * sizeof(T))); }
template <class T>
static void wrv(std::ofstream& f, const T* p, size_t n) { f.write((const char*)p, (std::streamsize)(n * sizeof(T))); }

static LinProb load(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
  char magic[8];
  rdv(f, magic, 8);
  if (memcmp(magic, "B200LIN1", 8)) { fprintf(stderr, "bad magic in %s\n", path.c_str()); exit(2); }
  LinProb p;
  rdv(f, &p.nvars, 1);
  p.var_dim.resize(p.nvars); rdv(f, p.var_dim.data(), p.nvars);
  p.ordering.resize(p.nvars); rdv(f, p.ordering.data(), p.nvars);
  int64_t ng;
  rdv(f, &ng, 1);
  p.groups.resize(ng);
  for (auto& g : p.groups) {
    rdv(f, &g.rows, 1); rdv(f, &g.arity, 1);
    g.dims.resize(g.arity); rdv(f, g.dims.data(), g.arity);
    rdv(f, &g.count, 1); rdv(f, &g.gi0, 1); rdv(f, &g.flags, 1);
    g.keys.resize(g.count * g.arity); rdv(f, g.keys.data(), g.keys.size());
    g.Ab.resize((size_t)g.count * g.rows * g.ncols()); rdv(f, g.Ab.data(), g.Ab.size());
    if (g.flags & 1) { g.sigmas.resize((size_t)g.count * g.rows); rdv(f, g.sigmas.data(), g.sigmas.size()); }
    if (g.flags & 4) { g.gidx.resize(g.count); rdv(f, g.gidx.data(), g.count); }
  }
  return p;
}

static void save(const LinProb& p, const std::string& path) {
  std::ofstream f(path, std::ios::binary);
  f.write("B200LIN1", 8);
  wrv(f, &p.nvars, 1);
  wrv(f, p.var_dim.data(), p.var_dim.size());
  wrv(f, p.ordering.data(), p.ordering.size());
  const int64_t ng = (int64_t)p.groups.size();
  wrv(f, &ng, 1);
  for (auto& g : p.groups) {
    wrv(f, &g.rows, 1); wrv(f, &g.arity, 1); wrv(f, g.dims.data(), g.dims.size());
    wrv(f, &g.count, 1); wrv(f, &g.gi0, 1); wrv(f, &g.flags, 1);
    wrv(f, g.keys.data(), g.keys.size()); wrv(f, g.Ab.data(), g.Ab.size());
    if (g.flags & 1) wrv(f, g.sigmas.data(), g.sigmas.size());
    if (g.flags & 4) wrv(f, g.gidx.data(), g.gidx.size());
  }
}

/* the graph in graph-position order, each factor a JacobianFactor(terms, b, Diagonal::Sigmas | unit) */
static gtsam::GaussianFactorGraph build_graph(const LinProb& p) {
  using namespace gtsam;
  std::vector<GaussianFactor::shared_ptr> slots(p.nfactors());
  for (auto& g : p.groups) {
    const int nc = g.ncols();
    for (int64_t i = 0; i < g.count; i++) {
      const double* M = g.Ab.data() + (size_t)i * g.rows * nc;
      const int64_t ppos = (g.flags & 4) ? g.gidx[i] : g.gi0 + i;
      if (g.flags & 8) {   // HessianFactor(keys, dims, SymmetricBlockMatrix-like augmented information)
        KeyVector keys;
        std::vector<DenseIndex> dims;
        for (int a = 0; a < g.arity; a++) { keys.push_back(Key(g.keys[i * g.arity + a])); dims.push_back(g.dims[a]); }
        dims.push_back(1);
        Matrix full(nc, nc);
        for (int c = 0; c < nc; c++) for (int r = 0; r < nc; r++) full(r, c) = r <= c ? M[r + (size_t)c * nc] : M[c + (size_t)r * nc];
        slots[ppos] = std::make_shared<HessianFactor>(keys, SymmetricBlockMatrix(dims, full));
        continue;
      }
      std::vector<std::pair<Key, Matrix>> terms;
      int col = 0;
      for (int a = 0; a < g.arity; a++) {
        Matrix A(g.rows, g.dims[a]);
        for (int c = 0; c < g.dims[a]; c++, col++) for (int r = 0; r < g.rows; r++) A(r, c) = M[r + (size_t)col * g.rows];
        terms.emplace_back(Key(g.keys[i * g.arity + a]), A);
      }
      Vector b(g.rows);
      for (int r = 0; r < g.rows; r++) b(r) = M[r + (size_t)col * g.rows];
      SharedDiagonal model;
      if (g.flags & 1) {
        Vector s(g.rows);
        for (int r = 0; r < g.rows; r++) s(r) = g.sigmas[(size_t)i * g.rows + r];
        model = noiseModel::Diagonal::Sigmas(s, /*smart=*/false);
      }
      const int64_t pos = (g.flags & 4) ? g.gidx[i] : g.gi0 + i;
      slots[pos] = std::make_shared<JacobianFactor>(terms, b, model);
    }
  }
  GaussianFactorGraph gfg;
  for (auto& f : slots) gfg.push_back(f);
  return gfg;
}

static gtsam::Ordering build_ordering(const LinProb& p) {
  gtsam::Ordering o;
  for (int64_t v : p.ordering) o.push_back(gtsam::Key(v));
  return o;
}

/* A GaussianFactorGraph of JacobianFactors -> groups by shape (rows, block widths, has-model), graph positions
 * explicit.  keys are mapped to dense ids in ascending Key order.  Returns false on a non-Jacobian or
 * constrained factor. */
static bool from_graph(const gtsam::GaussianFactorGraph& gfg, const gtsam::Ordering& ordering, LinProb* out,
                       std::map<gtsam::Key, int64_t>* id_of = nullptr) {
  using namespace gtsam;
  std::map<Key, int64_t> ids;
  std::map<Key, int> dimOf;
  for (auto& f : gfg) {
    if (!f) continue;
    if (!std::dynamic_pointer_cast<JacobianFactor>(f) && !std::dynamic_pointer_cast<HessianFactor>(f)) return false;
    for (auto it = f->begin(); it != f->end(); ++it) { ids[*it] = 0; dimOf[*it] = (int)f->getDim(it); }
  }
  int64_t n = 0;
  for (auto& kv : ids) kv.second = n++;
  out->nvars = n;
  out->var_dim.resize(n);
  for (auto& kv : ids) out->var_dim[kv.second] = dimOf[kv.first];
  out->ordering.clear();
  for (Key k : ordering) out->ordering.push_back(ids.at(k));
  std::map<std::vector<int>, size_t> sig2group;
  int64_t pos = 0;
  for (auto& f : gfg) {
    if (!f) continue;
    if (auto hf = std::dynamic_pointer_cast<HessianFactor>(f)) {
      std::vector<int> sig{-1, 0};
      for (auto it = hf->begin(); it != hf->end(); ++it) sig.push_back((int)hf->getDim(it));
      auto found = sig2group.find(sig);
      if (found == sig2group.end()) {
        JGroup g;
        g.arity = (int32_t)hf->size(); g.flags = 4 | 8; g.gi0 = -1;
        for (auto it = hf->begin(); it != hf->end(); ++it) g.dims.push_back((int32_t)hf->getDim(it));
        g.rows = g.ncols();
        out->groups.push_back(g);
        found = sig2group.emplace(sig, out->groups.size() - 1).first;
      }
      JGroup& g = out->groups[found->second];
      for (auto it = hf->begin(); it != hf->end(); ++it) g.keys.push_back(ids.at(*it));
      const Matrix info = hf->info().selfadjointView();
      for (int c = 0; c < info.cols(); c++) for (int r = 0; r < info.rows(); r++) g.Ab.push_back(info(r, c));
      g.gidx.push_back(pos++);
      g.count++;
      continue;
    }
    auto jf = std::dynamic_pointer_cast<JacobianFactor>(f);
    const SharedDiagonal& model = jf->get_model();
    if (model && model->isConstrained()) return false;
    const bool has_model = model && !model->isUnit();
    std::vector<int> sig{(int)jf->rows(), has_model ? 1 : 0};
    for (auto it = jf->begin(); it != jf->end(); ++it) sig.push_back((int)jf->getDim(it));
    auto found = sig2group.find(sig);
    if (found == sig2group.end()) {
      JGroup g;
      g.rows = (int32_t)jf->rows(); g.arity = (int32_t)jf->size(); g.flags = 4 | (has_model ? 1 : 0); g.gi0 = -1;
      for (auto it = jf->begin(); it != jf->end(); ++it) g.dims.push_back((int32_t)jf->getDim(it));
      out->groups.push_back(g);
      found = sig2group.emplace(sig, out->groups.size() - 1).first;
    }
    JGroup& g = out->groups[found->second];
    for (auto it = jf->begin(); it != jf->end(); ++it) g.keys.push_back(ids.at(*it));
    const Matrix Ab = jf->augmentedJacobianUnweighted();
    for (int c = 0; c < Ab.cols(); c++) for (int r = 0; r < Ab.rows(); r++) g.Ab.push_back(Ab(r, c));
    if (has_model) { const Vector s = model->sigmas(); for (int r = 0; r < s.size(); r++) g.sigmas.push_back(s(r)); }
    g.gidx.push_back(pos++);
    g.count++;
  }
  if (id_of) *id_of = ids;
  return true;
}

}  // namespace linio
