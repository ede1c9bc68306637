"""CPU, world_size 2, gloo: the N>1 path of the batch-sharded data parallelism (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from perm_equiv_graph_neural_cdes_b200 import dist as pdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        # a "batch of trajectories" and a tiny model whose loss is a sum over trajectories
        B = 5
        x = torch.arange(B * 3, dtype=torch.float32).reshape(B, 3)
        w = torch.nn.Parameter(torch.tensor([0.5, -1.0, 2.0]))
        b = torch.nn.Parameter(torch.tensor([0.1]))
        (xs,) = pdist.shard_batch([x], rank, world)
        loss = ((xs * w).sum(dim=1) + b).pow(2).sum()
        loss.backward()
        flat = pdist.allreduce_gradients([w, b])
        t = pdist.max_over_ranks(float(rank + 1), torch.device("cpu"))
        out[rank] = (flat.clone(), w.grad.clone(), t, tuple(xs.shape))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_batch_exactly():
    for total in (1, 5, 8, 50):
        for world in (1, 2, 3, 8):
            spans = [pdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pdist.shard_range(4, 2, 2)


def test_sharded_gradients_allreduce_to_the_full_batch_gradient():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    # reference: the whole batch on one process
    B = 5
    x = torch.arange(B * 3, dtype=torch.float32).reshape(B, 3)
    w = torch.nn.Parameter(torch.tensor([0.5, -1.0, 2.0]))
    b = torch.nn.Parameter(torch.tensor([0.1]))
    ((x * w).sum(dim=1) + b).pow(2).sum().backward()
    ref = torch.cat([w.grad, b.grad])
    for r in range(world):
        flat, wgrad, t, shape = out[r]
        assert torch.allclose(flat, ref)
        assert torch.allclose(wgrad, w.grad)
        assert t == 2.0
    assert out[0][3] == (3, 3) and out[1][3] == (2, 3)


# ---------------------------------------------------------------------------------------------------
# row-sharded mode: the partition logic, world size 2 over gloo
# ---------------------------------------------------------------------------------------------------
def _rowshard_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from perm_equiv_graph_neural_cdes_b200.rowshard import row_range

        g = torch.Generator().manual_seed(7)
        n, d = 256, 8
        A, Ad = torch.rand((n, n), generator=g, dtype=torch.float64), torch.rand((n, n), generator=g, dtype=torch.float64)
        M = torch.randn((n, d), generator=g, dtype=torch.float64)
        al, be, ga, de = 1.1, 0.9, 0.05, -0.03
        r0, r1 = row_range(n, rank, world)
        # what a rank holds: its rows of the path and of the TRANSPOSED path, and its rows of M
        E_rows = al * A[r0:r1] + be * Ad[r0:r1]
        Gt_rows = (ga * A[:, r0:r1] + de * Ad[:, r0:r1]).t()
        M_rows = M[r0:r1].contiguous()
        gathered = [torch.empty_like(M_rows) for _ in range(world)]
        dist.all_gather(gathered, M_rows)                       # the exchange step: every rank needs all rows of the operand
        M_full = torch.cat(gathered)
        colsum = M_rows.sum(0)
        dist.all_reduce(colsum)                                 # ... and the column sums over all nodes
        out_rows = M_rows + E_rows @ M_full + Gt_rows @ M_full  # no reduce-scatter: the transposed strip is local
        out[rank] = (r0, r1, out_rows, colsum)
    finally:
        dist.destroy_process_group()


def test_row_ranges_are_equal_strips_of_128_row_blocks():
    from perm_equiv_graph_neural_cdes_b200.rowshard import row_range

    for n, world in ((256, 2), (1024, 4), (16384, 8), (128, 1)):
        spans = [row_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert all((e - b) % 128 == 0 and e - b == n // world for b, e in spans)
    with pytest.raises(ValueError):
        row_range(1000, 0, 2)


def test_row_sharded_layer_equals_the_full_layer():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rowshard_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    g = torch.Generator().manual_seed(7)
    n, d = 256, 8
    A, Ad = torch.rand((n, n), generator=g, dtype=torch.float64), torch.rand((n, n), generator=g, dtype=torch.float64)
    M = torch.randn((n, d), generator=g, dtype=torch.float64)
    full = M + (1.1 * A + 0.9 * Ad) @ M + (0.05 * A - 0.03 * Ad).t() @ M
    got = torch.cat([out[r][2] for r in range(world)])
    assert torch.allclose(got, full, rtol=1e-12, atol=1e-12)
    assert torch.allclose(out[0][3], M.sum(0)) and torch.allclose(out[1][3], M.sum(0))
