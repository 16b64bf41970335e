"""CPU tests (gloo, world_size 2) of the multi-GPU host logic in torch_nfft_b200/dist.py.

The compute stages are injected: a numpy stand-in engine built from the oracle's stage functions
replaces the C-ABI stages, so what is tested here is the sharding plan and the collective
(partial oversampled grids summed over ranks; batch entries owned by ranks)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nfft_oracle as O
from torch_nfft_b200 import dist as D


class NumpyEngine:
    """Oracle-backed stand-in for dist.CudaEngine (same method signatures, CPU tensors)."""

    def spread(self, x, pos, batch, B, N, m):
        n, d = pos.shape
        x2 = x.numpy().reshape(n, -1)
        b = np.zeros(n, dtype=np.int64) if batch is None else batch.numpy()
        g = O.spread(pos.numpy(), x2, b, B, N, m)
        g = g if np.iscomplexobj(x2) else g.real
        return torch.from_numpy(np.ascontiguousarray(g.reshape((B * x2.shape[1],) + (2 * N,) * d))).to(
            torch.complex64 if np.iscomplexobj(x2) else torch.float32)

    def adjoint_finish(self, grid, d, B, cols, N, m, real_output):
        g = grid.numpy().reshape((B, grid.shape[0] // B) + (2 * N,) * d)
        return torch.from_numpy(O.adjoint_finish(g.astype(np.complex128), N, m, cols, real_output))

    def forward_begin(self, xhat, d, m, real_output):
        g = O.forward_begin(xhat.numpy(), d, m)
        g = g.reshape((-1,) + g.shape[2:])
        return torch.from_numpy(np.ascontiguousarray(g.real if real_output else g))

    def gather(self, grid, pos, batch, B, cols, N, m):
        n, d = pos.shape
        g = grid.numpy().reshape((B, grid.shape[0] // B) + (2 * N,) * d)
        b = np.zeros(n, dtype=np.int64) if batch is None else batch.numpy()
        y = O.gather(g.astype(np.complex64), pos.numpy(), b, N, m).reshape((n,) + tuple(cols))
        return torch.from_numpy(y.astype(np.complex64) if grid.is_complex() else y.real.astype(np.float32))

    def fastsum_middle(self, grid, coeffs, d, B, N, m):
        g = grid.numpy().reshape((B, grid.shape[0] // B) + (2 * N,) * d)
        g2 = O.fastsum_middle(g.astype(np.complex128), coeffs.numpy(), m)
        g2 = g2.reshape(grid.shape)
        return torch.from_numpy(np.ascontiguousarray(g2 if grid.is_complex() else g2.real.astype(np.float32)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    rng = np.random.default_rng(0)
    d, N, m, B, n = 2, 16, 3, 3, 120
    pos = rng.random((n * B, d), dtype=np.float32) - 0.5
    batch = np.repeat(np.arange(B, dtype=np.int64), n)
    x = rng.standard_normal((n * B, 2)).astype(np.float32)
    return d, N, m, B, pos, batch, x


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d, N, m, B, pos, batch, x = _problem()
        eng = NumpyEngine()
        tp, tb, tx = torch.from_numpy(pos), torch.from_numpy(batch), torch.from_numpy(x)
        out = {}
        # ---- point sharding: each rank holds a contiguous slice of the points of ALL batch entries
        lo, hi = D.shard_points(pos.shape[0], world, rank)
        y = D.nfft_adjoint_point_sharded(tx[lo:hi], tp[lo:hi], tb[lo:hi], N, m, batch_size=B, engine=eng)
        out["adj_point"] = O.rel_l2(y.numpy(), O.nfft_adjoint(x, pos, batch, N, m))
        full = O.nfft_adjoint(x, pos, batch, N, m)
        f = D.nfft_forward_point_sharded(torch.from_numpy(full), tp[lo:hi], tb[lo:hi], m, real_output=True, engine=eng)
        out["fwd_point"] = O.rel_l2(f.numpy(), O.nfft_forward(full, pos, batch, m, real_output=True)[lo:hi])
        co = O.gaussian_analytic_coeffs(0.2, d, N)
        s = D.nfft_fastsum_point_sharded(tx[lo:hi], torch.from_numpy(co), tp[lo:hi] * 0.5, source_batch=tb[lo:hi],
                                         cutoff=m, batch_size=B, engine=eng)
        ref = O.nfft_fastsum(x, co, (pos * 0.5).astype(np.float32), None, batch, batch, m=m)
        out["fastsum_point"] = O.rel_l2(s.numpy(), ref[lo:hi])
        # ---- point sharding with B a multiple of the world size: reduce-scatter into whole grids per rank
        n2 = 2 * (pos.shape[0] // B)                       # the points of batch entries 0 and 1
        pos2, batch2, x2 = pos[:n2], batch[:n2], x[:n2]
        lo2, hi2 = D.shard_points(n2, world, rank)
        full2 = O.nfft_adjoint(x2, pos2, batch2, N, m)
        y2 = D.nfft_adjoint_point_sharded(tx[lo2:hi2], tp[lo2:hi2], tb[lo2:hi2], N, m, batch_size=2, engine=eng)
        out["adj_point_rs"] = O.rel_l2(y2.numpy(), full2)
        y2l, (o_lo, o_hi) = D.nfft_adjoint_point_sharded(tx[lo2:hi2], tp[lo2:hi2], tb[lo2:hi2], N, m, batch_size=2,
                                                         engine=eng, scatter_output=True)
        out["adj_point_rs_local"] = O.rel_l2(y2l.numpy(), full2[o_lo:o_hi])
        out["owned_rs"] = (o_lo, o_hi)
        # ---- batch sharding: each rank transforms the batch entries it owns
        adj = lambda xx, pp, bb, NN, mm, ro, batch_size: torch.from_numpy(
            O.nfft_adjoint(xx.numpy(), pp.numpy(), bb.numpy(), NN, mm, ro) if pp.shape[0] else
            np.zeros((batch_size,) + (NN,) * d + (2,), np.complex64))
        yb, (b_lo, b_hi) = D.nfft_adjoint_batch_sharded(tx, tp, tb, N, m, batch_size=B, adjoint_fn=adj)
        out["adj_batch"] = O.rel_l2(yb.numpy(), full[b_lo:b_hi]) if b_hi > b_lo else 0.0
        yall, _ = D.nfft_adjoint_batch_sharded(tx, tp, tb, N, m, batch_size=B, adjoint_fn=adj, gather_output=True)
        out["adj_batch_gathered"] = O.rel_l2(yall.numpy(), full)
        fwd = lambda xh, pp, bb, mm, ro, batch_size: torch.from_numpy(O.nfft_forward(xh.numpy(), pp.numpy(), bb.numpy(), mm, ro))
        fb, (p_lo, p_hi) = D.nfft_forward_batch_sharded(torch.from_numpy(full), tp, tb, m, True, forward_fn=fwd)
        out["fwd_batch"] = O.rel_l2(fb.numpy(), O.nfft_forward(full, pos, batch, m, real_output=True)[p_lo:p_hi])
        out["owned"] = (b_lo, b_hi, p_lo, p_hi)
        results[rank] = out
    finally:
        dist.destroy_process_group()


def test_sharded_transforms_world_size_2():
    world = 2
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
        res = dict(results)
    assert set(res) == {0, 1}
    for rank, out in res.items():
        for key in ("adj_point", "adj_point_rs", "adj_point_rs_local", "fwd_point", "fastsum_point", "adj_batch",
                    "adj_batch_gathered", "fwd_batch"):
            assert out[key] < 1e-5, (rank, key, out[key])
    assert res[0]["owned_rs"] == (0, 1) and res[1]["owned_rs"] == (1, 2)
    # ownership: disjoint, contiguous, covering
    assert res[0]["owned"][:2] == (0, 2) and res[1]["owned"][:2] == (2, 3)
    assert res[0]["owned"][3] == res[1]["owned"][2] and res[1]["owned"][3] == 360


def test_shard_plans():
    assert [D.split_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [D.split_range(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    batch = torch.tensor([0, 0, 0, 1, 1, 3, 3, 3, 3])  # entry 2 is empty
    plans = [D.shard_batches(batch, 4, 2, r) for r in range(2)]
    assert plans == [(0, 2, 0, 5), (2, 4, 5, 9)]
    covered = sum(p[3] - p[2] for p in plans)
    assert covered == batch.numel()
