// symbolic.cpp — see symbolic.h.
#include "symbolic.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <thread>

namespace b200 {

AmalgOptions amalg_options_from_env() {
  AmalgOptions o;
  if (getenv("B200_NO_AMALGAMATE")) o.tol = -1.0;
  if (const char* e = getenv("B200_AMALG_TOL")) o.tol = atof(e);
  if (const char* e = getenv("B200_AMALG_SMALL")) o.small = atoi(e);
  if (const char* e = getenv("B200_AMALG_THIN_F")) o.thin_f = atoi(e);
  if (const char* e = getenv("B200_AMALG_THIN_TOL")) o.thin_tol = atof(e);
  return o;
}

bool build_symbolic(int64_t n, const int* var_dim, const int64_t* ordering, int64_t m,
                    const int64_t* fptr, const int64_t* fkeys, Symbolic* S, const char** err,
                    const AmalgOptions& amalg) {
  const bool timing = getenv("B200_SETUP_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[b200 symbolic] %-26s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  S->nvars = n;
  S->var_dim.assign(var_dim, var_dim + n);
  S->var_dof.assign(n + 1, 0);
  for (int64_t v = 0; v < n; v++) S->var_dof[v + 1] = S->var_dof[v] + var_dim[v];
  std::vector<int64_t> pos(n, -1);
  for (int64_t j = 0; j < n; j++) {
    const int64_t v = ordering[j];
    if (v < 0 || v >= n || pos[v] != -1) { *err = "ordering is not a permutation of the variable ids"; return false; }
    pos[v] = j;
  }
  // ---- variable index (CSR, factor positions ascending per variable) ----------
  std::vector<int64_t> vi_ptr(n + 1, 0);
  for (int64_t i = 0; i < m; i++) {
    if (fptr[i + 1] <= fptr[i]) { *err = "factor without keys"; return false; }
    for (int64_t q = fptr[i]; q < fptr[i + 1]; q++) {
      if (fkeys[q] < 0 || fkeys[q] >= n) { *err = "factor key out of range"; return false; }
      for (int64_t r = fptr[i]; r < q; r++) if (fkeys[r] == fkeys[q]) { *err = "factor lists a key twice"; return false; }
      vi_ptr[fkeys[q] + 1]++;
    }
  }
  for (int64_t v = 0; v < n; v++) vi_ptr[v + 1] += vi_ptr[v];
  std::vector<int64_t> vi(vi_ptr[n]);
  {
    std::vector<int64_t> cur(vi_ptr.begin(), vi_ptr.end() - 1);
    for (int64_t i = 0; i < m; i++)
      for (int64_t q = fptr[i]; q < fptr[i + 1]; q++) vi[cur[fkeys[q]]++] = i;
  }
  lap("variable index");
  // ---- elimination tree with path compression ----------------------------------
  std::vector<int64_t> eparent(n, -1), anc(n, -1), prevCol(m, -1), node_of_factor(m, -1);
  std::vector<int64_t> first_child(n, -1), last_child(n, -1), next_sib(n, -1);
  for (int64_t j = 0; j < n; j++) {
    const int64_t v = ordering[j];
    for (int64_t q = vi_ptr[v]; q < vi_ptr[v + 1]; q++) {
      const int64_t i = vi[q];
      if (prevCol[i] != -1) {
        int64_t r = prevCol[i];
        while (anc[r] != -1 && anc[r] != j) {
          const int64_t nx = anc[r];
          anc[r] = j;
          r = nx;
        }
        if (anc[r] == -1 && r != j) {
          anc[r] = j;
          eparent[r] = j;
          if (last_child[j] == -1) first_child[j] = r; else next_sib[last_child[j]] = r;
          last_child[j] = r;
        }
      } else {
        node_of_factor[i] = j;
      }
      prevCol[i] = j;
    }
  }
  lap("elimination tree");
  // node -> own factors (CSR, ascending factor position)
  std::vector<int64_t> nf_ptr(n + 1, 0);
  for (int64_t i = 0; i < m; i++) nf_ptr[node_of_factor[i] + 1]++;
  for (int64_t j = 0; j < n; j++) nf_ptr[j + 1] += nf_ptr[j];
  std::vector<int64_t> nfac(m);
  {
    std::vector<int64_t> cur(nf_ptr.begin(), nf_ptr.end() - 1);
    for (int64_t i = 0; i < m; i++) nfac[cur[node_of_factor[i]]++] = i;
  }
  lap("node factor lists");
  // ---- symbolic elimination: separator (as positions) of every etree node -------
  std::vector<int64_t> sep_off(n + 1, 0), sep_pool;
  sep_pool.reserve((size_t)(fptr[m] + n));
  std::vector<int64_t> mark(n, -1);
  std::vector<std::vector<int64_t>> cfront(n);  // frontal positions of the cluster headed by j
  std::vector<char> alive(n, 1);
  std::vector<int64_t> absorbed_into(n, -1);
  for (int64_t j = 0; j < n; j++) {
    const size_t s0 = sep_pool.size();
    for (int64_t q = nf_ptr[j]; q < nf_ptr[j + 1]; q++) {
      const int64_t i = nfac[q];
      for (int64_t a = fptr[i]; a < fptr[i + 1]; a++) {
        const int64_t pj = pos[fkeys[a]];
        if (pj != j && mark[pj] != j) { mark[pj] = j; sep_pool.push_back(pj); }
      }
    }
    for (int64_t c = first_child[j]; c != -1; c = next_sib[c])
      for (int64_t q = sep_off[c]; q < sep_off[c + 1]; q++) {
        const int64_t pj = sep_pool[q];
        if (pj != j && mark[pj] != j) { mark[pj] = j; sep_pool.push_back(pj); }
      }
    sep_off[j + 1] = (int64_t)sep_pool.size();
    const int64_t myNrParents = (int64_t)(sep_pool.size() - s0);
    // merge rule of JunctionTree-inst.h:98-118 + Cluster::mergeChildren
    int64_t myNrFrontals = 1;
    std::vector<int64_t>& fr = cfront[j];
    fr.push_back(j);
    bool merged_any = false;
    for (int64_t c = first_child[j]; c != -1; c = next_sib[c]) {
      const int64_t childParents = sep_off[c + 1] - sep_off[c];
      if (myNrParents + myNrFrontals == childParents) {
        myNrFrontals += (int64_t)cfront[c].size();
        fr.insert(fr.end(), cfront[c].rbegin(), cfront[c].rend());
        std::vector<int64_t>().swap(cfront[c]);
        alive[c] = 0;
        absorbed_into[c] = j;
        merged_any = true;
      }
    }
    if (merged_any) std::reverse(fr.begin(), fr.end());
  }
  lap("symbolic elimination");
  // ---- the reference's cliques, kept for reporting (Symbolic::ref) ---------------------
  {
    RefCliques& R = S->ref;
    std::vector<int64_t> rid(n, -1);
    int64_t rc = 0;
    for (int64_t j = 0; j < n; j++) if (alive[j]) rid[j] = rc++;
    std::vector<int64_t> rclique_of_node(n, -1);
    for (int64_t j = n - 1; j >= 0; j--) rclique_of_node[j] = alive[j] ? rid[j] : rclique_of_node[absorbed_into[j]];
    R.ncliques = rc;
    R.front_ptr.assign(rc + 1, 0); R.sep_ptr.assign(rc + 1, 0); R.parent.assign(rc, -1);
    R.nf.assign(rc, 0); R.ns.assign(rc, 0); R.level.assign(rc, 0); R.super.assign(rc, -1);
    R.front_vars.reserve(n);
    for (int64_t j = 0; j < n; j++) {
      if (!alive[j]) continue;
      const int64_t c = rid[j];
      int fd = 0, sd = 0;
      for (int64_t pj : cfront[j]) { R.front_vars.push_back(ordering[pj]); fd += var_dim[ordering[pj]]; }
      R.front_ptr[c + 1] = (int64_t)R.front_vars.size();
      const size_t s0 = R.sep_vars.size();
      for (int64_t q = sep_off[j]; q < sep_off[j + 1]; q++) { R.sep_vars.push_back(ordering[sep_pool[q]]); sd += var_dim[ordering[sep_pool[q]]]; }
      std::sort(R.sep_vars.begin() + (int64_t)s0, R.sep_vars.end());   // gtsam/linear/Scatter.cpp:69-72
      R.sep_ptr[c + 1] = (int64_t)R.sep_vars.size();
      R.nf[c] = fd; R.ns[c] = sd;
      if (eparent[j] != -1) R.parent[c] = rclique_of_node[eparent[j]];
      R.max_nf = std::max<int64_t>(R.max_nf, fd); R.max_ns = std::max<int64_t>(R.max_ns, sd);
      R.flops += (double)fd * fd * fd / 3.0 + (double)fd * fd * sd + (double)fd * sd * sd;
    }
    int maxlvl = 0;
    for (int64_t c = 0; c < rc; c++) {
      if (R.parent[c] >= 0) R.level[R.parent[c]] = std::max(R.level[R.parent[c]], R.level[c] + 1);
      maxlvl = std::max(maxlvl, R.level[c]);
    }
    R.nlevels = rc ? maxlvl + 1 : 0;
    // ---- relaxed amalgamation into supernodes (symbolic.h) -------------------------------
    if (amalg.tol >= 0.0) {
      std::vector<int64_t> head(rc, -1);            // head node (last frontal position) of every reference clique
      for (int64_t j = 0; j < n; j++) if (alive[j]) head[rid[j]] = j;
      std::vector<int64_t> into(rc), F(rc), Sd(rc);
      for (int64_t c = 0; c < rc; c++) { into[c] = c; F[c] = R.nf[c]; Sd[c] = R.ns[c]; }
      auto find = [&](int64_t c) { while (into[c] != c) c = into[c] = into[into[c]]; return c; };
      bool any = false;
      for (int64_t c = 0; c < rc; c++) {             // children have smaller ids than parents
        if (R.parent[c] < 0) continue;
        if (R.level[c] == 0 && R.nf[c] <= amalg.leaf_max_f) continue;   // leaves keep their own (fused) kernels
        const int64_t p = find(R.parent[c]);
        const int64_t width = F[p] + Sd[p], fill = width - Sd[c];
        if ((double)fill <= amalg.tol * (double)width || F[c] + Sd[c] < amalg.small ||
            (F[c] <= amalg.thin_f && (double)fill <= amalg.thin_tol * (double)width)) {
          into[c] = p; F[p] += F[c]; any = true;
        }
      }
      if (any) {
        for (int64_t c = 0; c < rc; c++) {
          const int64_t r = find(c);
          if (r == c) continue;
          std::vector<int64_t>& src = cfront[head[c]];
          std::vector<int64_t>& dst = cfront[head[r]];
          dst.insert(dst.end(), src.begin(), src.end());
          std::vector<int64_t>().swap(src);
          alive[head[c]] = 0;
        }
        for (int64_t c = 0; c < rc; c++) if (find(c) == c) std::sort(cfront[head[c]].begin(), cfront[head[c]].end());
        // every node follows its clique's head to the head of the supernode
        for (int64_t j = 0; j < n; j++) {
          const int64_t c = rclique_of_node[j], r = find(c);
          if (r != c || !alive[j]) { if (j != head[r]) absorbed_into[j] = head[r]; }
        }
      }
      for (int64_t c = 0; c < rc; c++) R.super[c] = (int)find(c);   // reference-clique id of the surviving head, renumbered below
    } else {
      for (int64_t c = 0; c < rc; c++) R.super[c] = (int)c;
    }
    // renumber R.super to supernode ids (ascending head position among the survivors)
    {
      std::vector<int> newid(rc, -1);
      int k = 0;
      for (int64_t c = 0; c < rc; c++) if (R.super[c] == (int)c) newid[c] = k++;
      for (int64_t c = 0; c < rc; c++) R.super[c] = newid[R.super[c]];
    }
  }
  lap("reference cliques + amalgamation");
  // ---- clique tables (supernodes) ------------------------------------------------------
  std::vector<int64_t> cid(n, -1);
  int64_t nc = 0;
  for (int64_t j = 0; j < n; j++) if (alive[j]) cid[j] = nc++;
  // node -> clique: follow absorption upwards (heads have larger positions)
  std::vector<int64_t> clique_of_node(n, -1);
  for (int64_t j = n - 1; j >= 0; j--) clique_of_node[j] = alive[j] ? cid[j] : clique_of_node[absorbed_into[j]];
  S->ncliques = nc;
  S->front_ptr.assign(nc + 1, 0); S->sep_ptr.assign(nc + 1, 0);
  S->parent.assign(nc, -1); S->nf.assign(nc, 0); S->ns.assign(nc, 0);
  S->off.assign(nc + 1, 0); S->level.assign(nc, 0);
  S->var_clique.assign(n, -1); S->var_slot.assign(n, -1);
  S->front_vars.clear(); S->sep_vars.clear();
  S->front_vars.reserve(n);
  for (int64_t j = 0; j < n; j++) {
    if (!alive[j]) continue;
    const int64_t c = cid[j];
    int slot = 0;
    for (int64_t pj : cfront[j]) {
      const int64_t v = ordering[pj];
      S->front_vars.push_back(v);
      S->var_clique[v] = (int)c;
      S->var_slot[v] = slot;
      slot += var_dim[v];
    }
    S->nf[c] = slot;
    S->front_ptr[c + 1] = (int64_t)S->front_vars.size();
    const size_t s0 = S->sep_vars.size();
    int sdim = 0;
    for (int64_t q = sep_off[j]; q < sep_off[j + 1]; q++) {
      const int64_t v = ordering[sep_pool[q]];
      S->sep_vars.push_back(v);
      sdim += var_dim[v];
    }
    // separator keys sorted by id (== by Key): gtsam/linear/Scatter.cpp:69-72
    std::sort(S->sep_vars.begin() + (int64_t)s0, S->sep_vars.end());
    S->ns[c] = sdim;
    S->sep_ptr[c + 1] = (int64_t)S->sep_vars.size();
    if (eparent[j] != -1) S->parent[c] = clique_of_node[eparent[j]];
    const int64_t nn = (int64_t)S->nf[c] + sdim + 1;
    S->off[c + 1] = S->off[c] + nn * nn;
    S->max_nf = std::max<int64_t>(S->max_nf, S->nf[c]);
    S->max_ns = std::max<int64_t>(S->max_ns, sdim);
    const double f = S->nf[c], s = sdim;
    S->flops += f * f * f / 3.0 + f * f * s + f * s * s;
  }
  S->arena_doubles = S->off[nc];
  // levels (children have smaller ids than parents)
  int maxlvl = 0;
  for (int64_t c = 0; c < nc; c++) {
    const int64_t p = S->parent[c];
    if (p >= 0) S->level[p] = std::max(S->level[p], S->level[c] + 1);
    maxlvl = std::max(maxlvl, S->level[c]);
  }
  S->nlevels = nc ? maxlvl + 1 : 0;
  S->lvl_ptr.assign(S->nlevels + 1, 0);
  for (int64_t c = 0; c < nc; c++) S->lvl_ptr[S->level[c] + 1]++;
  for (int64_t l = 0; l < S->nlevels; l++) S->lvl_ptr[l + 1] += S->lvl_ptr[l];
  S->lvl_cliques.assign(nc, 0);
  {
    std::vector<int64_t> cur(S->lvl_ptr.begin(), S->lvl_ptr.end() - 1);
    for (int64_t c = 0; c < nc; c++) S->lvl_cliques[cur[S->level[c]]++] = (int)c;
  }
  lap("supernode tables");
  // ---- scatter maps ------------------------------------------------------------------
  // factors by owning clique
  S->fac_clique.assign(m, -1); S->fac_slot0.assign(m, -1); S->fac_slot1.assign(m, -1);
  S->fac_slots.assign(fptr[m], -1);
  std::vector<int64_t> cf_ptr(nc + 1, 0);
  for (int64_t i = 0; i < m; i++) {
    S->fac_clique[i] = (int)clique_of_node[node_of_factor[i]];
    cf_ptr[S->fac_clique[i] + 1]++;
  }
  for (int64_t c = 0; c < nc; c++) cf_ptr[c + 1] += cf_ptr[c];
  std::vector<int64_t> cf(m);
  {
    std::vector<int64_t> cur(cf_ptr.begin(), cf_ptr.end() - 1);
    for (int64_t i = 0; i < m; i++) cf[cur[S->fac_clique[i]]++] = i;
  }
  // children by clique
  std::vector<int64_t> ch_ptr(nc + 1, 0);
  for (int64_t c = 0; c < nc; c++) if (S->parent[c] >= 0) ch_ptr[S->parent[c] + 1]++;
  for (int64_t c = 0; c < nc; c++) ch_ptr[c + 1] += ch_ptr[c];
  std::vector<int64_t> ch(ch_ptr[nc]);
  {
    std::vector<int64_t> cur(ch_ptr.begin(), ch_ptr.end() - 1);
    for (int64_t c = 0; c < nc; c++) if (S->parent[c] >= 0) ch[cur[S->parent[c]]++] = c;
  }
  S->ea_ptr.assign(nc + 1, 0); S->didx_ptr.assign(nc + 1, 0);
  for (int64_t c = 0; c < nc; c++) {
    S->ea_ptr[c + 1] = S->ea_ptr[c] + S->ns[c] + 1;
    S->didx_ptr[c + 1] = S->didx_ptr[c] + S->nf[c] + S->ns[c];
  }
  S->ea_map.assign(S->ea_ptr[nc], -1);
  S->didx.assign(S->didx_ptr[nc], -1);
  // Every clique writes only its own ranges (its didx rows, the slots of the factors it owns, the ea_map of its
  // children), so the cliques are processed in parallel: chunks handed out through an atomic counter, one slot[]
  // scratch per thread.  (4.5 of the 7.3 s of this phase at 10M factors were spent here on one thread.)
  lap("scatter prelude");
  std::atomic<int64_t> next_chunk(0);
  std::atomic<int> failure(0);   // 1: factor variable missing, 2: child separator variable missing
  const int64_t chunk = 2048;
  auto worker = [&]() {
    std::vector<int> slot(n, -1);
    for (;;) {
      const int64_t c0 = next_chunk.fetch_add(chunk);
      if (c0 >= nc || failure.load()) return;
      const int64_t c1 = std::min(nc, c0 + chunk);
      for (int64_t c = c0; c < c1; c++) {
        int k = 0;
        int64_t dq = S->didx_ptr[c];
        for (int64_t q = S->front_ptr[c]; q < S->front_ptr[c + 1]; q++) {
          const int64_t v = S->front_vars[q];
          slot[v] = k;
          for (int t = 0; t < var_dim[v]; t++) S->didx[dq++] = (int)(S->var_dof[v] + t);
          k += var_dim[v];
        }
        for (int64_t q = S->sep_ptr[c]; q < S->sep_ptr[c + 1]; q++) {
          const int64_t v = S->sep_vars[q];
          slot[v] = k;
          for (int t = 0; t < var_dim[v]; t++) S->didx[dq++] = (int)(S->var_dof[v] + t);
          k += var_dim[v];
        }
        const int nn = k + 1;
        for (int64_t q = cf_ptr[c]; q < cf_ptr[c + 1]; q++) {
          const int64_t i = cf[q];
          for (int64_t a = fptr[i]; a < fptr[i + 1]; a++) {
            S->fac_slots[a] = slot[fkeys[a]];
            if (S->fac_slots[a] < 0) failure.store(1);
          }
          S->fac_slot0[i] = S->fac_slots[fptr[i]];
          if (fptr[i + 1] - fptr[i] >= 2) S->fac_slot1[i] = S->fac_slots[fptr[i] + 1];
        }
        for (int64_t q = ch_ptr[c]; q < ch_ptr[c + 1]; q++) {
          const int64_t cc = ch[q];
          int64_t e = S->ea_ptr[cc];
          for (int64_t qq = S->sep_ptr[cc]; qq < S->sep_ptr[cc + 1]; qq++) {
            const int64_t v = S->sep_vars[qq];
            if (slot[v] < 0) { failure.store(2); continue; }
            for (int t = 0; t < var_dim[v]; t++) S->ea_map[e++] = slot[v] + t;
          }
          S->ea_map[e < S->ea_ptr[cc + 1] ? e : S->ea_ptr[cc + 1] - 1] = nn - 1;
        }
        // reset scratch
        for (int64_t q = S->front_ptr[c]; q < S->front_ptr[c + 1]; q++) slot[S->front_vars[q]] = -1;
        for (int64_t q = S->sep_ptr[c]; q < S->sep_ptr[c + 1]; q++) slot[S->sep_vars[q]] = -1;
      }
    }
  };
  {
    int nthreads = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), std::min<int64_t>(16, (nc + chunk - 1) / chunk));
    if (const char* e = getenv("B200_SYMBOLIC_THREADS")) nthreads = std::max(1, atoi(e));
    if (timing) fprintf(stderr, "[b200 symbolic] scatter maps on %d threads (hardware_concurrency %u, %lld cliques)\n", nthreads, std::thread::hardware_concurrency(), (long long)nc);
    if (nthreads <= 1) {
      worker();
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < nthreads; t++) pool.emplace_back(worker);
      for (auto& t : pool) t.join();
    }
  }
  lap("scatter maps (threads)");
  if (failure.load() == 1) { *err = "internal: factor variable not in owning clique"; return false; }
  if (failure.load() == 2) { *err = "internal: child separator variable not in parent clique"; return false; }
  return true;
}

}  // namespace b200
