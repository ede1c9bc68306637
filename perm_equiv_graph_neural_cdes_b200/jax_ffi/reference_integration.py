"""Reference-side integration of the fused path (JAX environment only; NOT importable in this repository's image).

Copy this file and ``pegncde_jax.py`` next to the reference's ``src/`` tree.  Nothing in the reference is edited:

* the config switch -- ``VectorFieldCfg.build`` resolves ``getattr(vector_fields, self.name)``
  (src/configs/vector_field_configs.py:52), so registering the fused module under a new name makes
  ``name: FusedPermEquivGraphVectorField`` in a YAML config select it (same constructor arguments, same leaves);
* the model wrapper -- a subclass of the reference's ``PGTGraphNeuralCDE`` whose ``__call__`` differs from
  src/models/pgt_graph_neural_cde.py:78-136 in exactly two lines: the control objects are packed once
  (``fused_control`` instead of two ``diffrax.CubicInterpolation``) and ``diffrax.diffeqsolve`` becomes ``fused_diffeqsolve``
  with the SAME keyword arguments.  ``jax.vmap(model)`` (src/configs/loss_configs.py:44), ``eqx.filter_jit`` and
  ``eqx.filter_value_and_grad`` (src/engine/trainer_pgt.py:346) work on it unchanged.
"""
import diffrax
import jax
import jax.numpy as jnp

import pegncde_jax as fused
from src.models import vector_fields
from src.models.pgt_graph_neural_cde import PGTGraphNeuralCDE

# --- config switch ------------------------------------------------------------------------------------------------
vector_fields.FusedPermEquivGraphVectorField = fused.FusedPermEquivGraphVectorField


# --- model wrapper ------------------------------------------------------------------------------------------------
class FusedPGTGraphNeuralCDE(PGTGraphNeuralCDE):
    def __call__(self, ts, coeffs_adj, x_coeffs, x0, evolving_out=False, global_readout=True):
        control = fused.fused_control(ts, coeffs_adj, x_coeffs, self.cfg.hidden_dim, len(self.vector_field.gnn_layers))
        term = diffrax.ODETerm(self.wrapped_vector_field)
        y0 = jax.vmap(self.encoder)(x0)
        saveat = diffrax.SaveAt(ts=ts) if evolving_out else diffrax.SaveAt(t1=True)
        latent_node_path = fused.fused_diffeqsolve(terms=term, solver=self.method, t0=ts[0], t1=ts[-1], dt0=0.1, y0=y0,
                                                   args=[control, None], stepsize_controller=self.controller, saveat=saveat)
        output = jax.vmap(self.decoder)(latent_node_path.ys[-1])
        return jnp.sum(output, axis=0) if global_readout else output
