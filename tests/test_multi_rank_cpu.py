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
            assert (co == -1).sum() >= 1          # a replicated top exists
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
