"""CPU, world_size 2 over gloo: the sharding plan of SURVEY §8(e) — every factor is owned by
exactly one rank, every fused leaf clique by one rank, and the per-rank partial sums the
multi-GPU path all-reduces (graph error, Hessian diagonal, shared top fronts expressed as the
dense normal-equation blocks of the top variables) add up to the single-rank result.  The
numeric kernels need a GPU; here each rank evaluates its shard with the C oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gtsam_b200 import capi, datasets, problem as P
from oracle import oracle_py as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _shard(prob, factor_owner, rank):
    groups = []
    for g in prob.groups:
        own = factor_owner[g.graph_index0:g.graph_index0 + g.count] == rank
        if not own.any():
            continue
        noise = g.noise.reshape(g.count, -1)[own] if g.noise_per_factor else g.noise
        groups.append(P.FactorGroup(g.type, g.keys[own], g.meas[own], g.noise_kind, noise,
                                    None if g.cal_index is None else g.cal_index[own]))
    return P.Problem(prob.var_type, prob.values, prob.ordering, groups, prob.cal)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = datasets.make("bal_tiny", ncams=12, npoints=300)
    co, fo = capi.shard_plan(prob, world)
    full = O.OracleProblem(prob)
    full.linearize()
    mine = O.OracleProblem(_shard(prob, fo, rank))
    mine.linearize()
    e = torch.tensor([mine.error()], dtype=torch.float64)
    h = torch.from_numpy(mine.hessian_diagonal().copy())
    dist.all_reduce(e)
    dist.all_reduce(h)
    ok_err = abs(e.item() - full.error()) <= 1e-12 * full.error()
    ok_hd = np.abs(h.numpy() - full.hessian_diagonal()).max() <= 1e-12 * np.abs(full.hessian_diagonal()).max()
    nown = torch.tensor([int((fo == rank).sum())])
    dist.all_reduce(nown)
    out[rank] = (ok_err, ok_hd, int(nown.item()) == prob.nfactors, int((co == rank).sum()))
    dist.destroy_process_group()


def test_shard_plan_partitions_everything(built):
    """BAL (Schur): the top is the camera chain, the subtrees are points; Pose3 sphere (nested-
    dissection-like natural ordering): the top is the upper separators, subtrees are branches."""
    for prob in (datasets.make("bal_tiny", ncams=12, npoints=300), datasets.make("sphere_tiny", layers=10, per_ring=16, ordering="reverse")):
        for world in (1, 2, 4, 8):
            co, fo = capi.shard_plan(prob, world)
            assert fo.min() >= 0 and fo.max() < world
            assert set(np.unique(co)) <= set(range(-1, world))
            if world == 1:
                assert np.all(co == 0) and np.all(fo == 0)
                continue
            # (a top need not exist: the scattered-visibility BAL generator yields two independent camera components,
            #  which two ranks take whole; from 4 ranks on the camera supernodes are the top and the points the subtrees)
            assert world == 2 or (co == -1).sum() >= 1
            # the top is ancestor-closed and every subtree lives on one rank
            from oracle import oracle_py as Oq
            par = Oq.OracleProblem(prob).cliques()[4]
            for c in range(len(par)):
                if par[c] >= 0:
                    assert co[par[c]] == -1 or co[par[c]] == co[c]
                    assert not (co[c] == -1 and co[par[c]] != -1)
            counts = np.bincount(fo, minlength=world)
            if prob.meta.get("kind") == "bal":
                # (rank 0 additionally owns the factors of the replicated top)
                assert counts[1:].max(initial=counts[0]) - counts[1:].min(initial=counts[0]) <= 0.15 * prob.nfactors + 12


def test_two_rank_partial_sums_over_gloo(built):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        ok_err, ok_hd, covered, nleaf = out[r]
        assert ok_err and ok_hd and covered and nleaf > 0


@pytest.mark.parametrize("world", [4, 8])
def test_shard_plan_of_a_nested_dissection_bal_graph(built, world):
    """bal_1m_metis (the reference's METIS ordering): the few heavy nested-dissection branches make the planner fall
    back to longest-processing-time-first packing; the plan stays valid (ancestor-closed top, one rank per subtree,
    every factor owned once) and every rank gets a comparable share of the factors."""
    import ctypes as C
    prob = datasets.make("bal_1m_metis")
    co, fo = capi.shard_plan(prob, world)
    L = capi.lib()
    desc, keep = prob.c_desc()
    h = C.c_void_p()
    capi._check(L.b200_symbolic_create(C.byref(desc), C.byref(h)))
    info = P.CSymbolicInfo()
    L.b200_symbolic_get_info(h, C.byref(info))
    fp, sp = np.zeros(info.ncliques + 1, dtype=np.int64), np.zeros(info.ncliques + 1, dtype=np.int64)
    fv, sv = np.zeros(info.frontal_list_len, dtype=np.int64), np.zeros(max(1, info.separator_list_len), dtype=np.int64)
    par = np.zeros(info.ncliques, dtype=np.int64)
    L.b200_symbolic_get_cliques(h, capi._ip(fp), capi._ip(fv), capi._ip(sp), capi._ip(sv), capi._ip(par))
    L.b200_symbolic_destroy(h)
    has_par = par >= 0
    pc = co[par[has_par]]
    cc = co[has_par]
    assert np.all((pc == -1) | (pc == cc))            # a non-top parent owns its whole subtree
    assert not np.any((cc == -1) & (pc != -1))        # the top is ancestor-closed
    assert (co == -1).sum() >= 1 and set(np.unique(fo)) == set(range(world))
    counts = np.bincount(fo, minlength=world)
    # the top is kept shallow (subtrees up to total / world: every top level is a communication stage of the distributed
    # top — measured at 8 GPUs on the 10M-factor graph: 3.68 ms per iteration against 4.57 with a 4x finer split), so the
    # nested-dissection branches are packed as they come; rank 0 also owns the factors of the top
    assert counts.min() > 0.6 * counts.max()
