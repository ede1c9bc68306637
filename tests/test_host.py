"""CPU: host logic of the product package and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import perm_equiv_graph_neural_cdes_b200 as P
from perm_equiv_graph_neural_cdes_b200 import _lib
from oracle import reference_path as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pegncde.h")).read()
    declared = set(re.findall(r"\b(pegncde_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    l = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(l, name), name


def test_strerror_and_version():
    l = _lib.lib()
    assert b"sm_100a" in l.pegncde_version()
    assert l.pegncde_strerror(0) == b"ok"
    assert b"workspace" in l.pegncde_strerror(3)
    assert b"unknown" in l.pegncde_strerror(99)


@pytest.mark.parametrize("n,h,e,L", [(129, 64, 8, 3), (1000, 64, 16, 3), (100, 32, 3, 3), (400, 16, 0, 2)])
def test_param_count_matches_reference_pytree(n, h, e, L):
    d = _lib.PegDims(1, n, (n + 31) // 32 * 32, h, e, L, 4, 0)
    widths = R.layer_widths(h, L, e, e > 0)
    expect = sum(widths[i + 1] * widths[i] + widths[i + 1] + 2 * widths[i] + 16 for i in range(L))
    assert _lib.lib().pegncde_param_count(d) == expect
    vf = P.PermEquivGraphVectorField(h, h, widths[-1], L, e, n, key=0)
    assert vf.flat_params().numel() == expect
    off = (ctypes.c_int64 * (5 * L))()
    assert _lib.lib().pegncde_param_offsets(d, off) == 0
    assert off[0] == 0 and off[4] == widths[1] * widths[0] + widths[1] + 2 * widths[0]


def test_bad_dims_are_rejected_without_touching_the_gpu():
    l = _lib.lib()
    good = dict(B=1, n=10, ldn=32, h=8, e=0, L=2, T=4, flags=0)
    for k, v in [("ldn", 10), ("ldn", 12), ("ldn", 64), ("h", 6), ("h", 512), ("L", 0), ("L", 9), ("T", 1), ("B", 0), ("n", 0)]:
        d = _lib.PegDims(**{**good, k: v})
        assert l.pegncde_param_count(d) == 0, (k, v)
        assert l.pegncde_workspace_bytes(d, 0, 1) == 0
    assert l.pegncde_workspace_bytes(_lib.PegDims(**good), 99, 1) == 0
    # a compute entry point validates before any CUDA call
    rc = l.pegncde_solve_fwd(None, _lib.PegDims(**{**good, "h": 6}), _lib.PegControl(), None, None, 1, None, None, None, None, None, 0)
    assert rc == 1
    rc = l.pegncde_solve_fwd(None, _lib.PegDims(**good), _lib.PegControl(), None, None, 1, None, None, None, None, None, 0)
    assert rc == 2  # null pointers


def test_workspace_grows_with_problem():
    l = _lib.lib()
    a = l.pegncde_workspace_bytes(_lib.PegDims(1, 100, 128, 32, 3, 3, 12, 0), _lib.PEG_WS_SOLVE_BWD, 10)
    b = l.pegncde_workspace_bytes(_lib.PegDims(4, 100, 128, 32, 3, 3, 12, 0), _lib.PEG_WS_SOLVE_BWD, 10)
    assert 0 < a < b


def test_step_table_matches_oracle():
    for rule in ("state", "prev_diff"):
        for (t1, dt) in [(3, 0.1), (1, 0.1), (1, 0.01), (8, 0.1), (2, 0.01), (5, 0.3)]:
            assert np.array_equal(P.constant_step_table(0.0, t1, dt, rule), R.constant_step_table(0.0, t1, dt, rule))


def test_hermite_coefficients_match_oracle():
    ts = torch.tensor([0.0, 0.5, 1.5, 2.0])
    ys = torch.randn(4, 5, 5, 2, generator=torch.Generator().manual_seed(0))
    for a, b in zip(P.backward_hermite_coefficients(ts, ys), R.backward_hermite_coefficients(ts, ys)):
        assert torch.equal(a, b)


def test_flat_params_layout_and_grad_routing():
    vf = P.PermEquivGraphVectorField(8, 8, 8, 2, 0, 5, key=3)
    flat = vf.flat_params()
    l0 = vf.gnn_layers[0]
    assert torch.equal(flat[:64], l0.conv_layer.linear.weight.reshape(-1))
    assert torch.equal(flat[64 + 8 + 16: 64 + 8 + 16 + 2], l0.param1)
    (flat * torch.arange(flat.numel())).sum().backward()
    assert torch.equal(l0.param2.grad, torch.tensor([64 + 8 + 16 + 2.0, 64 + 8 + 16 + 3.0]))


def test_cpu_tensors_are_refused():
    vf = P.PermEquivGraphVectorField(8, 8, 8, 2, 0, 5, key=3)
    ts = torch.arange(4.0)
    A = torch.rand(4, 5, 5)
    co = R.reference_layout_coeffs(ts, A)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.diffeqsolve(P.ODETerm(vf), P.Tsit5(), 0.0, 3.0, 0.1, torch.zeros(5, 8), P.CubicInterpolation(ts, co))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "perm_equiv_graph_neural_cdes_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("load_oracle_layers", ""), os.path.join(dirpath, f)


def test_tiled_layout_is_a_permutation_with_contiguous_warp_loads():
    """The 32x32-tiled plane layout: a bijection onto [0, 4*npad^2), 16-KB tiles, and every (g, plane, m) warp
    row of 32 lanes x 4 floats is 512 contiguous bytes."""
    from perm_equiv_graph_neural_cdes_b200.control import tiled_offsets

    npad = 96
    off = tiled_offsets(npad)
    assert sorted(off.reshape(-1).tolist()) == list(range(4 * npad * npad))
    tile = off[:, 32:64, 64:96]  # tile (1, 2)
    assert int(tile.min()) == (1 * 3 + 2) * 4096 and int(tile.max()) == (1 * 3 + 2) * 4096 + 4095
    # elements of one micro-tile row are 4 consecutive floats
    assert torch.equal(off[2, 5, 8:12] - off[2, 5, 8], torch.arange(4))
