"""Solve wrappers with the reference's call signatures, running the fused solve.

``PGTGraphNeuralCDE``  <- src/models/pgt_graph_neural_cde.py:13-136
``TGBGraphNeuralCDE``  <- src/models/tgb_graph_neural_cde.py:13-171 (learned node-signal control)
``GraphNeuralCDE``     <- src/models/graph_neural_cde.py:13-113 (PID-controlled adaptive solve + dense output)

Only the ``diffeqsolve`` call is replaced; encoder / decoder MLPs are ordinary device ops
(host plumbing in the reference too).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .control import CubicInterpolation, LinearInterpolation, backward_hermite_coefficients
from .solve import ConstantStepSize, ODETerm, PIDController, SaveAt, Tsit5, diffeqsolve
from .vector_field import CDEWrapperVectorField, Linear, PermEquivGraphVectorField


class MLP(nn.Module):
    """``eqx.nn.MLP(in, out, width_size, depth)``: ``depth`` hidden layers, ReLU, identity output."""

    def __init__(self, in_size: int, out_size: int, width_size: int, depth: int, generator=None):
        super().__init__()
        sizes = [in_size] + [width_size] * depth + [out_size]
        self.layers = nn.ModuleList([Linear(sizes[i], sizes[i + 1], generator) for i in range(len(sizes) - 1)])

    def forward(self, x):
        for i, lin in enumerate(self.layers):
            x = torch.nn.functional.linear(x, lin.weight, lin.bias)
            if i < len(self.layers) - 1:
                x = torch.relu(x)
        return x


class PGTGraphNeuralCDE(nn.Module):
    """model(ts, coeffs_adj, x_coeffs, x0, evolving_out=False, global_readout=True)."""

    def __init__(self, hidden_dim: int, data_dim: int, feature_dim: int, vector_field: PermEquivGraphVectorField,
                 interpolation: str = "cubic", seed: int = 0, dt0: float = 0.1):
        super().__init__()
        if interpolation not in ("cubic", "linear"):
            raise ValueError("interpolation must be 'cubic' or 'linear' (pgt_graph_neural_cde.py:101-107)")
        g = torch.Generator().manual_seed(seed)
        self.hidden_dim = hidden_dim
        self.interpolation = interpolation
        self.vector_field = vector_field
        self.encoder = MLP(data_dim, hidden_dim, 16, 2, g)
        self.decoder = MLP(hidden_dim, feature_dim, 16, 2, g)
        self.method = Tsit5()
        self.controller = ConstantStepSize()
        self.wrapped_vector_field = CDEWrapperVectorField(vector_field, hidden_dim)
        self.dt0 = dt0

    def forward(self, ts, coeffs_adj, x_coeffs, x0, evolving_out: bool = False, global_readout: bool = True):
        interp = LinearInterpolation if self.interpolation == "linear" else CubicInterpolation
        control_adj = coeffs_adj if not isinstance(coeffs_adj, (tuple, list, torch.Tensor)) else interp(ts, coeffs_adj)
        control_data = x_coeffs if not isinstance(x_coeffs, (tuple, list, torch.Tensor)) else interp(ts, x_coeffs)
        y0 = self.encoder(x0)
        ts_host = ts.detach().reshape(-1).cpu()     # one host copy: t0, t1 and the save times are host scalars of the C-ABI
        saveat = SaveAt(ts=ts_host) if evolving_out else SaveAt(t1=True)   # pgt_graph_neural_cde.py:114-117
        sol = diffeqsolve(terms=ODETerm(self.wrapped_vector_field), solver=self.method, t0=float(ts_host[0]),
                          t1=float(ts_host[-1]), dt0=self.dt0, y0=y0, args=[control_adj, control_data],
                          stepsize_controller=self.controller, saveat=saveat)
        output = self.decoder(sol.ys[-1])
        if global_readout:
            return output.sum(dim=-2)
        return output


class TGBGraphNeuralCDE(nn.Module):
    """``model(ts, coeffs_adj, x_data, x0, start_time, evolving_out=False)`` of src/models/tgb_graph_neural_cde.py:96-171.

    The node-signal control is LEARNED: ``x_data [T, n, num_nodes]`` goes through ``data_encoder`` (Linear
    num_nodes -> data_embed_dim), is stacked with the time channel and turned into Hermite coefficients inside the model
    (``:118-137``), so reverse mode continues from the solve (``g_xcoef`` of ``pegncde_solve_bwd``) into the encoder.
    ConstantStepSize, ``dt0 = 0.01`` (``:143``); ``SaveAt(t1=True)``, or ``SaveAt(ts=ts)`` with ``evolving_out=True`` (``:147-150``)."""

    def __init__(self, hidden_dim: int, vector_field: PermEquivGraphVectorField, use_mlps: bool = True, seed: int = 0,
                 dt0: float = 0.01, return_sequence: bool = False):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        n, e = vector_field.num_nodes, vector_field.data_embed_dim
        self.hidden_dim, self.dt0, self.return_sequence = hidden_dim, dt0, return_sequence
        self.vector_field = vector_field
        if use_mlps:
            self.encoder, self.decoder = MLP(n, hidden_dim, 16, 2, g), MLP(hidden_dim, n, 16, 2, g)
        else:
            self.encoder, self.decoder = Linear(n, hidden_dim, g), Linear(hidden_dim, n, g)
        self.data_encoder = Linear(n, e, g)
        self.method, self.controller = Tsit5(), ConstantStepSize()
        self.wrapped_vector_field = CDEWrapperVectorField(vector_field, hidden_dim)

    @staticmethod
    def _lin(mod, x):
        return mod(x) if isinstance(mod, MLP) else torch.nn.functional.linear(x, mod.weight, mod.bias)

    def forward(self, ts, coeffs_adj, x_data, x0, start_time=None, evolving_out: bool = False):
        x_emb = self._lin(self.data_encoder, x_data)                                  # [T, n, e]
        tsf = ts.to(x_emb.dtype)
        x_path = torch.stack([tsf[:, None, None].expand_as(x_emb), x_emb], dim=-1)      # [T, n, e, 2] (time, value)
        coeffs_data = backward_hermite_coefficients(tsf, x_path)
        control_adj = coeffs_adj if not isinstance(coeffs_adj, (tuple, list, torch.Tensor)) else CubicInterpolation(ts, coeffs_adj)
        control_data = CubicInterpolation(ts, coeffs_data)
        y0 = self._lin(self.encoder, x0)
        ts_host = ts.detach().reshape(-1).cpu()
        saveat = SaveAt(ts=ts_host) if evolving_out else SaveAt(t1=True)   # tgb_graph_neural_cde.py:147-150
        sol = diffeqsolve(terms=ODETerm(self.wrapped_vector_field), solver=self.method, t0=float(ts_host[0]), t1=float(ts_host[-1]),
                          dt0=self.dt0, y0=y0, args=[control_adj, control_data], stepsize_controller=self.controller,
                          saveat=saveat)
        return self._lin(self.decoder, sol.ys if self.return_sequence else sol.ys[-1])


class GraphNeuralCDE(nn.Module):
    """``model(ts, coeffs_adj, x0, evolving_out=True)`` of src/models/graph_neural_cde.py:60-113: Linear(1->h) encoder,
    ``ODETerm(vector_field)`` without the CDE wrapper, ``dt0=None``, ``PIDController(rtol=1e-3, atol=1e-6)``,
    ``SaveAt(ts=ts)`` (dense output) or ``SaveAt(t1=True)``, Linear(h->1) read-out.

    ``controller=ConstantStepSize()`` + ``dt0`` selects the fixed-step fused solve instead (one C-ABI call; read-out at
    every step boundary) -- not a reference configuration, kept for benchmarking."""

    def __init__(self, hidden_dim: int, vector_field: PermEquivGraphVectorField, seed: int = 0, dt0: Optional[float] = None,
                 return_sequence: bool = True, controller=None):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.initial_linear = Linear(1, hidden_dim, g)
        self.final_linear = Linear(hidden_dim, 1, g)
        self.vector_field = vector_field
        self.dt0 = dt0
        self.return_sequence = return_sequence
        self.method = Tsit5()
        self.controller = controller if controller is not None else PIDController(rtol=1e-3, atol=1e-6)

    def forward(self, ts, coeffs_adj, x0, evolving_out: bool = True):
        control_adj = coeffs_adj if not isinstance(coeffs_adj, (tuple, list, torch.Tensor)) else CubicInterpolation(ts, coeffs_adj)
        y0 = torch.nn.functional.linear(x0, self.initial_linear.weight, self.initial_linear.bias)
        tsv = ts.reshape(-1, ts.shape[-1])[0]   # a batch shares one time grid (dataset_configs.py:160-170)
        if isinstance(self.controller, PIDController):
            saveat = SaveAt(ts=tsv) if evolving_out else SaveAt(t1=True)
        else:
            saveat = SaveAt(steps=evolving_out, t1=not evolving_out)
        sol = diffeqsolve(terms=ODETerm(self.vector_field), solver=self.method, t0=float(tsv[0]), t1=float(tsv[-1]),
                          dt0=self.dt0, y0=y0, args=control_adj, stepsize_controller=self.controller, saveat=saveat)
        ys = sol.ys if self.return_sequence else sol.ys[-1]
        return torch.nn.functional.linear(ys, self.final_linear.weight, self.final_linear.bias)
