// symbolic.h — host-side symbolic phase: variable index -> elimination tree ->
// junction tree (supernodes) -> level schedule + scatter maps.
//
// Produces exactly the cliques the reference builds on every solve
// (gtsam/inference/VariableIndex-inl.h:27-50,
//  gtsam/inference/EliminationTree-inst.h:77-155,
//  gtsam/inference/JunctionTree-inst.h:63-151,
//  gtsam/inference/ClusterTree-inst.h:46-95) — including its child-merge rule,
// under which at most the first qualifying child is absorbed — but computes
// them once per problem, with path-compressed root finding, and emits flat
// integer tables the device kernels walk level by level.
#pragma once
#include <cstdint>
#include <vector>

namespace b200 {

// The cliques exactly as the reference's JunctionTree builds them (reporting only: b200_get_cliques,
// b200_get_conditional, the shim's Bayes tree).  The numeric tables of `Symbolic` below describe the
// SUPERNODES the device eliminates: the same cliques after relaxed amalgamation (see build_symbolic).
struct RefCliques {
  int64_t ncliques = 0, nlevels = 0, max_nf = 0, max_ns = 0;
  std::vector<int64_t> front_ptr, front_vars, sep_ptr, sep_vars, parent;
  std::vector<int> nf, ns, level;
  std::vector<int> super;             // supernode (numeric clique id) that holds each reference clique
  double flops = 0;
};

struct Symbolic {
  RefCliques ref;
  int64_t nvars = 0, ncliques = 0, nlevels = 0;
  std::vector<int> var_dim;
  std::vector<int64_t> var_dof;       // nvars+1 prefix of dims (delta layout)
  // cliques, numbered by ascending elimination position of their head variable
  std::vector<int64_t> front_ptr, front_vars, sep_ptr, sep_vars;
  std::vector<int64_t> parent;        // -1 for roots
  std::vector<int> nf, ns;            // scalar frontal / separator dims
  std::vector<int64_t> off;           // offset (doubles) of the (nf+ns+1)^2 col-major front
  std::vector<int> level;
  std::vector<int64_t> lvl_ptr;       // nlevels+1
  std::vector<int> lvl_cliques;       // cliques grouped by level (ascending id inside a level)
  std::vector<int64_t> ea_ptr;        // ncliques+1; per clique ns+1 entries
  std::vector<int> ea_map;            // scalar row/col in the parent front of each trailing index
  std::vector<int64_t> didx_ptr;      // ncliques+1; per clique nf+ns entries
  std::vector<int> didx;              // index into delta of each front row/col
  std::vector<int> var_clique, var_slot;  // owning clique + scalar slot of each variable
  std::vector<int> fac_clique;        // per graph position: owning clique
  std::vector<int> fac_slot0, fac_slot1;  // scalar slots of key 0 / key 1 in that clique (-1 if unary)
  std::vector<int> fac_slots;         // the same for every key of every factor, laid out like fkeys (n-ary factors)
  int64_t arena_doubles = 0;
  int64_t max_nf = 0, max_ns = 0;
  double flops = 0;
};

// Relaxed supernode amalgamation (numeric tree only; the reference's cliques are kept in Symbolic::ref):
// a non-leaf clique c is merged into its parent p when the explicit zeros this adds to c's rows,
// (f_p + s_p - s_c) columns, are at most `amalg_tol` of the parent's width f_p + s_p, or when c is tiny
// (f_c + s_c < amalg_small), or when c is thin (f_c <= thin_f pivots: a whole front life cycle — load, one pivot step,
// extend-add, a level of the tree — for a handful of pivots) and the zeros stay below thin_tol of the width.  sep(c) is a subset of front(p) U sep(p), so the merged supernode has frontals
// F_c U F_p (in elimination order) and the parent's separator; structurally-zero entries stay exactly 0.0
// through the dense partial Cholesky, so every conditional [R S d] of the reference's cliques can be read
// back out of the supernode.  amalg_tol < 0 disables (supernodes == reference cliques).  Camera chains of
// BAL graphs (6-pivot cliques with 600-1200 separator columns, one per level) collapse into their parents:
// bal_c4_metis 25 -> 8 levels, 899 -> 279 MB of fronts, +5 % flops.
struct AmalgOptions { double tol = 0.15; int small = 48; int leaf_max_f = 6; int thin_f = 24; double thin_tol = 0.5; };
AmalgOptions amalg_options_from_env();

// fptr (nfactors+1) / fkeys: CSR of the variable ids of every factor by graph position, in the
// factor's own key order (any arity >= 1, as a JacobianFactor has).  Returns false (and fills err)
// on invalid input.
bool build_symbolic(int64_t nvars, const int* var_dim, const int64_t* ordering, int64_t nfactors,
                    const int64_t* fptr, const int64_t* fkeys, Symbolic* out, const char** err,
                    const AmalgOptions& amalg = amalg_options_from_env());

}  // namespace b200
