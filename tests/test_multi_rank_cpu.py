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
    prob = datasets.make("bal_tiny", ncams=12, npoints=300)
    for world in (1, 2, 4, 8):
        co, fo = capi.shard_plan(prob, world)
        assert fo.min() >= 0 and fo.max() < world
        assert set(np.unique(co)) <= set(range(-1, world))
        # balanced by factor count over the fused leaf cliques (BAL points: 6 factors each)
        # (rank 0 additionally owns the factors of the replicated top: priors, and the factors of
        # the few points the reference's merge rule absorbs into a camera clique)
        counts = np.bincount(fo, minlength=world)
        assert counts[1:].max(initial=counts[0]) - counts[1:].min(initial=counts[0]) <= 0.05 * prob.nfactors + 12
        # the top (camera cliques) is replicated
        assert (co == -1).sum() >= 1


def test_two_rank_partial_sums_over_gloo(built):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        ok_err, ok_hd, covered, nleaf = out[r]
        assert ok_err and ok_hd and covered and nleaf > 0
