"""Control paths: the host-side mirror of ``diffrax.backward_hermite_coefficients`` /
``diffrax.CubicInterpolation`` as the reference uses them
(src/configs/dataset_configs.py:147-173, 1073-1100; src/models/pgt_graph_neural_cde.py:101-107),
plus the packed planar device layout the kernels consume.

Coefficient tuples keep diffrax's order ``(d, c, b, a)`` and the reference's layout
``[..., T-1, n, n, 2]`` with the last axis interleaved (time, adjacency).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from ._lib import PegControl, PegDims, check, lib


def backward_hermite_coefficients(ts: torch.Tensor, ys: torch.Tensor) -> Tuple[torch.Tensor, ...]:
    """``diffrax.backward_hermite_coefficients(ts, ys)`` -> ``(d, c, b, a)``, each ``[T-1, ...]``.

    Elementwise torch ops (runs on whatever device ``ys`` lives on).  Per interval i with
    dt = t[i+1]-t[i] and secant m_i: a = y_i, b = m_{i-1} (m_0 first), c = 2(m_i-b)/dt, d = -(m_i-b)/dt^2."""
    ts = ts.to(ys.dtype)
    dt = (ts[1:] - ts[:-1]).reshape((-1,) + (1,) * (ys.dim() - 1))
    m = (ys[1:] - ys[:-1]) / dt
    b = torch.cat([m[:1], m[:-1]], dim=0)
    a = ys[:-1]
    c = 2.0 * (m - b) / dt
    d = -(m - b) / (dt * dt)
    return d, c, b, a


def _pad32(n: int) -> int:
    return (n + 31) // 32 * 32


def tiled_offsets(npad: int) -> torch.Tensor:
    """Float offset, inside one slab, of element (plane q, row i, col k) of the 32x32-tiled plane layout
    (``peg_tile_off`` in csrc/peg_common.cuh) as an int64 tensor [4, npad, npad] -- for inspection / tests."""
    nt = npad // 32
    i = torch.arange(npad).view(1, npad, 1)
    k = torch.arange(npad).view(1, 1, npad)
    q = torch.arange(4).view(4, 1, 1)
    rt, ct, r, c = i // 32, k // 32, i % 32, k % 32
    rq, m, cq, e = r // 4, r % 4, c // 4, c % 4
    s = (rq - cq) % 8
    g, lane = s // 4, (s % 4) * 8 + cq
    return (rt * nt + ct) * 4096 + ((g * 4 + q) * 4 + m) * 128 + lane * 4 + e


class CubicInterpolation:
    """Drop-in for ``diffrax.CubicInterpolation(ts, coeffs)`` on the fused path.

    ``coeffs = (d, c, b, a)``; for the adjacency control each is ``[T-1, n, n, 2]`` (or batched
    ``[B, T-1, n, n, 2]``), for the node-signal control ``[T-1, n, e, 2]``.  ``evaluate`` /
    ``derivative`` exist for API parity (torch ops on the device); the solver never calls them --
    it consumes the packed planes built lazily by :meth:`packed_adj` / :meth:`packed_x`.
    """

    def __init__(self, ts: torch.Tensor, coeffs: Sequence[torch.Tensor]):
        if isinstance(coeffs, torch.Tensor):  # the PGT trainer passes the 4 arrays stacked (trainer_pgt.py:203)
            coeffs = tuple(coeffs[i] for i in range(4))
        self.ts = ts
        self.d, self.c, self.b, self.a = coeffs
        self._packed = None

    # -- diffrax API -------------------------------------------------------------------
    def _interpret_t(self, t):
        ts = self.ts.to(torch.float32)
        t = torch.as_tensor(t, dtype=torch.float32, device=ts.device)
        idx = torch.searchsorted(ts.contiguous(), t, right=False) - 1
        idx = int(idx.clamp(0, ts.numel() - 2))
        return idx, t - ts[idx]

    def evaluate(self, t):
        i, s = self._interpret_t(t)
        return self.a[i] + s * (self.b[i] + s * (self.c[i] + s * self.d[i]))

    def derivative(self, t):
        i, s = self._interpret_t(t)
        return self.b[i] + s * (2.0 * self.c[i] + 3.0 * s * self.d[i])


class LinearInterpolation(CubicInterpolation):
    """Drop-in for ``diffrax.LinearInterpolation(ts, ys)`` (``interpolation="linear"`` of the PGT / TGB models,
    src/models/pgt_graph_neural_cde.py:101-103): ``ys`` are the knot values ``[T, n, n, 2]`` (what ``diffrax.linear_interpolation``
    returns for data without NaNs).  A piecewise-linear path is the cubic path with ``a = y_i``, ``b = (y_{i+1} - y_i) / dt``,
    ``c = d = 0``, so the fused kernels run it unchanged (same left-continuous piece lookup)."""

    def __init__(self, ts: torch.Tensor, ys: torch.Tensor):
        tsv = ts.to(ys.dtype)
        batched = ys.dim() == 5 or (ys.dim() == 4 and ts.dim() == 2)
        t_axis = 1 if batched else 0
        dt = (tsv[..., 1:] - tsv[..., :-1])
        shape = [1] * ys.dim()
        shape[t_axis] = -1
        if tsv.dim() == 2:
            shape[0] = tsv.shape[0]
        a = ys.narrow(t_axis, 0, ys.shape[t_axis] - 1)
        b = (ys.narrow(t_axis, 1, ys.shape[t_axis] - 1) - a) / dt.reshape(shape)
        z = torch.zeros_like(a)
        super().__init__(ts, (z, z, b, a))


class PackedControl:
    """Planar device layout of one batch of control paths (``PegControl`` in pegncde.h).

    Built once per batch; owns the device tensors and hands raw pointers to the C-ABI."""

    def __init__(self, B: int, n: int, T: int, e: int, device):
        self.B, self.n, self.T, self.e = B, n, T, e
        self.ldn = _pad32(n)  # planes are stored as 32x32 tiles, zero padded
        f = dict(dtype=torch.float32, device=device)
        self.ts = torch.empty((B, T), **f)
        self.adj_coef = torch.empty((B, T - 1, 4 * self.ldn * self.ldn), **f)
        self.adj_rowsum = torch.empty((B, T - 1, 4, n), **f)
        self.adj_diag = torch.empty((B, T - 1, 4, n), **f)
        self.adj_total = torch.empty((B, T - 1, 4), **f)
        self.tch_coef = torch.empty((B, T - 1, 3, n), **f)
        self.adj_absmax = torch.empty((B, T - 1, 4), **f)   # max |entry| of each plane: range bound of the fp16x2 operand format
        self.x_coef = torch.empty((B, T - 1, 3, n, 2 * e), **f) if e > 0 else None
        self.x_packed = None   # differentiable source of x_coef when the node-signal coefficients require grad
        # state of the adjacency part, SHARED by every view made with with_x(): column sums of the planes ([B,T-1,4,n], built on
        # demand for the directed fusion layer), the host (d,c,b,a) arrays whose planes have not been copied / packed yet
        # (streamed controls) and the knot times shared by the batch (streaming needs one time grid; numpy fp32)
        self._adj = {"colsum": None, "pending": None, "host_ts": None}

    adj_colsum = property(lambda self: self._adj["colsum"], lambda self, v: self._adj.__setitem__("colsum", v))
    pending = property(lambda self: self._adj["pending"], lambda self, v: self._adj.__setitem__("pending", v))
    host_ts = property(lambda self: self._adj["host_ts"], lambda self, v: self._adj.__setitem__("host_ts", v))

    def with_x(self, x_coeffs) -> "PackedControl":
        """The same adjacency planes (shared tensors, shared streaming state -- no copy, no re-pack) paired with the
        node-signal coefficients ``x_coeffs = (d,c,b,a)``, each ``[T-1,n,e,2]`` or ``[B,T-1,n,e,2]``."""
        v = PackedControl.__new__(PackedControl)
        v.B, v.n, v.T, v.ldn = self.B, self.n, self.T, self.ldn
        for name in ("ts", "adj_coef", "adj_rowsum", "adj_diag", "adj_total", "tch_coef", "adj_absmax"):
            setattr(v, name, getattr(self, name))
        v._adj = self._adj
        v._keepalive = self
        _attach_x(v, x_coeffs, self.device)
        return v

    def pack_pieces(self, begin: int, count: int, staging, stream_ptr: int) -> None:
        """Packs cubic pieces [begin, begin+count) from device staging arrays (d,c,b,a each [B,count,n,n,2])."""
        check(lib().pegncde_pack_adj_range(stream_ptr, self.dims(h=4, L=1), begin, count, staging[0].data_ptr(), staging[1].data_ptr(),
                                           staging[2].data_ptr(), staging[3].data_ptr(), self.adj_coef.data_ptr(),
                                           self.adj_rowsum.data_ptr(), self.adj_diag.data_ptr(), self.adj_total.data_ptr(),
                                           self.tch_coef.data_ptr()), "pegncde_pack_adj_range")
        self.plane_maxima(begin, count, stream_ptr)

    def plane_maxima(self, begin: int = 0, count: int = -1, stream_ptr: Optional[int] = None) -> None:
        """``adj_absmax`` of the cubic pieces [begin, begin+count) from the tiled planes (``pegncde_adj_absmax``)."""
        count = self.T - 1 - begin if count < 0 else count
        st = _stream_ptr(self.device) if stream_ptr is None else stream_ptr
        check(lib().pegncde_adj_absmax(st, self.dims(h=4, L=1), begin, count, self.adj_coef.data_ptr(), self.adj_absmax.data_ptr()),
              "pegncde_adj_absmax")

    def materialize(self) -> "PackedControl":
        """Copies and packs every pending piece now (adaptive solves, single evaluations, ragged time grids)."""
        if self.pending is not None:
            dev = self.device
            total = 4 * self.pending[0].numel() * 4
            chunk = self.T - 1 if total <= (8 << 30) else max(1, (self.T - 1) * (8 << 30) // total)   # <= 8 GB of staging at a time
            for iv in range(0, self.T - 1, chunk):
                cnt = min(chunk, self.T - 1 - iv)
                staging = [c[:, iv:iv + cnt].to(dev, non_blocking=True).contiguous() for c in self.pending]
                self.pack_pieces(iv, cnt, staging, _stream_ptr(dev))
            self.pending = None
        return self

    def select(self, b: int) -> "PackedControl":
        """View of graph ``b`` as a batch of one (no copy): adaptive solves step every trajectory on its own."""
        v = PackedControl.__new__(PackedControl)
        v.B, v.n, v.T, v.e, v.ldn = 1, self.n, self.T, self.e, self.ldn
        for name in ("ts", "adj_coef", "adj_rowsum", "adj_diag", "adj_total", "tch_coef", "adj_absmax"):
            setattr(v, name, getattr(self, name)[b:b + 1])
        v.x_coef = self.x_coef[b:b + 1] if self.x_coef is not None else None
        v.x_packed = None
        cs = self.adj_colsum
        v._adj = {"colsum": cs[b:b + 1] if cs is not None else None, "pending": None, "host_ts": None}
        v._keepalive = self
        return v

    def dims(self, h: int, L: int, flags: int = 0) -> PegDims:
        return PegDims(self.B, self.n, self.ldn, h, self.e, L, self.T, flags)

    def dense_planes(self) -> torch.Tensor:
        """Un-tiles ``adj_coef`` to [B, T-1, 4(a,b,c,d), n, n] (inspection / tests)."""
        off = tiled_offsets(self.ldn).to(self.adj_coef.device)
        dense = self.adj_coef[:, :, off.reshape(-1)].reshape(self.B, self.T - 1, 4, self.ldn, self.ldn)
        return dense[..., : self.n, : self.n]

    def struct(self) -> PegControl:
        return PegControl(
            self.ts.data_ptr(), self.adj_coef.data_ptr(), self.adj_rowsum.data_ptr(), self.adj_diag.data_ptr(),
            self.adj_total.data_ptr(), self.tch_coef.data_ptr(), self.x_coef.data_ptr() if self.x_coef is not None else None,
            self.adj_colsum.data_ptr() if self.adj_colsum is not None else None,
            self.adj_absmax.data_ptr() if getattr(self, "adj_absmax", None) is not None else None,
        )

    def ensure_colsums(self) -> "PackedControl":
        """Column sums of the planes (``pegncde_adj_colsums``), needed by ``PEG_FLAG_DIRECTED`` only; computed once."""
        if self.adj_colsum is None:
            self.materialize()
            self.adj_colsum = torch.empty((self.B, self.T - 1, 4, self.n), dtype=torch.float32, device=self.device)
            check(lib().pegncde_adj_colsums(_stream_ptr(self.device), self.dims(h=4, L=1), self.adj_coef.data_ptr(),
                                            self.adj_colsum.data_ptr()), "pegncde_adj_colsums")
        return self

    @property
    def device(self):
        return self.ts.device


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _as_batched(x: torch.Tensor, nd_unbatched: int) -> torch.Tensor:
    return x.unsqueeze(0) if x.dim() == nd_unbatched else x


def _attach_x(pc: "PackedControl", x_coeffs, device) -> None:
    """Sets the node-signal part (e, x_coef, x_packed) of ``pc`` from reference-layout coefficients (or clears it)."""
    pc.x_packed = None
    if x_coeffs is None:
        pc.e, pc.x_coef = 0, None
        return
    if isinstance(x_coeffs, torch.Tensor):
        x_coeffs = tuple(x_coeffs[i] for i in range(4))
    cx = [_as_batched(c, 4).to(device=device, dtype=torch.float32, non_blocking=True).contiguous() for c in x_coeffs]
    B, Tm1, n, e = cx[0].shape[0], cx[0].shape[1], cx[0].shape[2], cx[0].shape[3]
    if (B, Tm1, n) != (pc.B, pc.T - 1, pc.n):
        raise ValueError(f"node-signal coefficients [B={B},T-1={Tm1},n={n}] do not match the adjacency control [B={pc.B},T-1={pc.T - 1},n={pc.n}]")
    pc.e = e
    if any(c.requires_grad for c in cx):
        # learnable node-signal path (TGB models, tgb_graph_neural_cde.py:118-137): the re-layout stays on the autograd
        # tape so that pegncde_solve_bwd's g_xcoef flows back into backward_hermite_coefficients / the data encoder
        pc.x_packed = torch.stack([cx[2], cx[1], cx[0]], dim=2).reshape(B, Tm1, 3, n, 2 * e)
        pc.x_coef = pc.x_packed.detach().contiguous()
    else:
        pc.x_coef = torch.empty((B, Tm1, 3, n, 2 * e), dtype=torch.float32, device=device)
        check(lib().pegncde_pack_x(_stream_ptr(device), pc.dims(h=4, L=1), cx[0].data_ptr(), cx[1].data_ptr(), cx[2].data_ptr(),
                                   cx[3].data_ptr(), pc.x_coef.data_ptr()), "pegncde_pack_x")
        pc._x_keepalive = cx   # sources stay alive until the pack kernel has run (stream-ordered)


def pack_control(
    ts: torch.Tensor,
    coeffs_adj: Sequence[torch.Tensor],
    x_coeffs: Optional[Sequence[torch.Tensor]] = None,
    device=None,
) -> PackedControl:
    """Reference-layout coefficients -> :class:`PackedControl` (runs ``pegncde_pack_adj`` / ``pegncde_pack_x``).

    ``coeffs_adj``: ``(d,c,b,a)`` each ``[T-1,n,n,2]`` or ``[B,T-1,n,n,2]`` (or the 4 stacked on axis 0);
    ``x_coeffs``: ``(d,c,b,a)`` each ``[T-1,n,e,2]`` or ``[B,T-1,n,e,2]``; ``ts``: ``[T]`` or ``[B,T]``."""
    if isinstance(coeffs_adj, torch.Tensor):
        coeffs_adj = tuple(coeffs_adj[i] for i in range(4))
    if isinstance(x_coeffs, torch.Tensor):
        x_coeffs = tuple(x_coeffs[i] for i in range(4))
    device = torch.device(device if device is not None else coeffs_adj[0].device)
    if device.type != "cuda":
        raise RuntimeError("pack_control needs a CUDA device: the fused path has no CPU fallback")
    if all(c.device.type == "cpu" for c in coeffs_adj):
        return _pack_control_streamed(ts, coeffs_adj, x_coeffs, device)
    cad = [_as_batched(c, 4).to(device=device, dtype=torch.float32).contiguous() for c in coeffs_adj]
    B, Tm1, n = cad[0].shape[0], cad[0].shape[1], cad[0].shape[2]
    pc = PackedControl(B, n, Tm1 + 1, 0, device)
    tsb = _as_batched(ts, 1).to(device=device, dtype=torch.float32)
    pc.ts.copy_(tsb.expand(B, Tm1 + 1))
    check(
        lib().pegncde_pack_adj(_stream_ptr(device), pc.dims(h=4, L=1), cad[0].data_ptr(), cad[1].data_ptr(), cad[2].data_ptr(),
                               cad[3].data_ptr(), pc.adj_coef.data_ptr(), pc.adj_rowsum.data_ptr(), pc.adj_diag.data_ptr(),
                               pc.adj_total.data_ptr(), pc.tch_coef.data_ptr()),
        "pegncde_pack_adj",
    )
    pc.plane_maxima()
    _attach_x(pc, x_coeffs, device)
    pc._keepalive = cad   # keep the sources alive until the pack kernels have run (stream-ordered)
    return pc


def _pack_control_streamed(ts, coeffs_adj, x_coeffs, device) -> PackedControl:
    """HOST coefficient arrays (what the reference's trainers hold: numpy / torch-CPU batches, trainer_pgt.py:201-207): the
    adjacency planes are NOT copied here.  The returned control is *pending*: a fixed-step solve copies and packs cubic piece
    i+1 (``pegncde_pack_adj_range``, side stream) while the steps inside piece i run; anything else calls
    :meth:`PackedControl.materialize` first.  The node-signal coefficients (small) are packed right away."""
    cad = []
    for c in coeffs_adj:
        c = _as_batched(c, 4).to(torch.float32).contiguous()
        cad.append(c if c.is_pinned() else c.pin_memory())
    B, Tm1, n = cad[0].shape[0], cad[0].shape[1], cad[0].shape[2]
    pc = PackedControl(B, n, Tm1 + 1, 0, device)
    tsb = _as_batched(ts, 1).to(torch.float32)
    pc.ts.copy_(tsb.expand(B, Tm1 + 1), non_blocking=True)
    _attach_x(pc, x_coeffs, device)
    pc.pending = cad
    grid = tsb.reshape(-1, Tm1 + 1)
    pc.host_ts = grid[0].cpu().numpy().copy() if bool((grid == grid[0]).all()) else None
    pc._keepalive = cad
    return pc


def build_control(ts: torch.Tensor, snapshots: torch.Tensor, x_t: Optional[torch.Tensor] = None, device=None) -> PackedControl:
    """Graph snapshots -> :class:`PackedControl` in one device pass (``pegncde_build_adj``): the fused replacement of
    ``get_graph_interpolation_coeffs`` (src/configs/dataset_configs.py:147-173, 1073-1100), which stacks a time channel
    onto ``A_k``, calls ``diffrax.backward_hermite_coefficients`` on the host and ships four ``[T-1,n,n,2]`` arrays.

    ``snapshots``: ``[T,n,n]`` or ``[B,T,n,n]``; ``ts``: ``[T]`` or ``[B,T]``; ``x_t`` (optional node signals
    ``[T,n,e]`` / ``[B,T,n,e]``): their coefficients are small and are built with :func:`backward_hermite_coefficients`."""
    device = torch.device(device if device is not None else snapshots.device)
    if device.type != "cuda":
        raise RuntimeError("build_control needs a CUDA device: the fused path has no CPU fallback")
    A = _as_batched(snapshots, 3).to(device=device, dtype=torch.float32).contiguous()
    B, T, n = A.shape[0], A.shape[1], A.shape[2]
    e = 0
    cx = None
    if x_t is not None:
        xb = _as_batched(x_t, 3).to(device=device, dtype=torch.float32)
        e = xb.shape[-1]
        tsx = _as_batched(ts, 1).to(device=device, dtype=torch.float32).expand(B, T)
        per = []
        for b in range(B):
            X = torch.stack([tsx[b][:, None, None].expand(T, n, e), xb[b]], dim=-1)
            per.append(backward_hermite_coefficients(tsx[b], X))
        cx = [torch.stack([per[b][i] for b in range(B)]).contiguous() for i in range(4)]
    pc = PackedControl(B, n, T, 0, device)
    pc.ts.copy_(_as_batched(ts, 1).to(device=device, dtype=torch.float32).expand(B, T))
    check(lib().pegncde_build_adj(_stream_ptr(device), pc.dims(h=4, L=1), pc.ts.data_ptr(), A.data_ptr(), pc.adj_coef.data_ptr(),
                                  pc.adj_rowsum.data_ptr(), pc.adj_diag.data_ptr(), pc.adj_total.data_ptr(), pc.tch_coef.data_ptr()),
          "pegncde_build_adj")
    pc.plane_maxima()
    _attach_x(pc, cx, device)
    pc._keepalive = A
    return pc
