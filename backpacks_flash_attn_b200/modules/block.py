"""Pre-norm residual `Block` (flash_attn/modules/block.py:22-106).

    mixer_out = mixer(h);  res = mixer_out + res;  h = norm1(res)
    mlp_out   = mlp(h);    res = mlp_out + res;    h = norm2(res)

With `fused_dropout_add_ln` each "add + LayerNorm" pair is one bp_ln_residual_fwd launch that keeps the
residual stream in fp32 (block.py:84-88, 101-105); without it the un-fused PyTorch sequence of the
reference is kept, including its rounding of the residual to the weight dtype before LayerNorm (:76).
Optionally (`fuse_residual_add`, off by default), when the branch ends in a FusedDense GEMM of this library (MHA.out_proj with fused_bias_fc,
FusedDenseGeluDense.fc2) and the residual stream is fp32 on the GPU, the "add" moves into that GEMM's epilogue
(bp_linear_bias_residual_fwd updates the residual stream in place) and the LayerNorm reads the fp32 residual
only (bp_ln_fwd): the 16-bit branch output is never written to HBM and read back, which halves the LayerNorm's
traffic.  The fp32 accumulator is added un-rounded, so this path is at least as accurate as the reference's.
Dropout and stochastic depth are identities in eval mode and not modelled; post-norm is out of scope.
"""
from __future__ import annotations

from functools import partial
from typing import Optional

import torch.nn as nn
from torch import Tensor

from ..ops.fused_dense import FusedDense, FusedDenseGeluDense, can_fuse_residual
from ..ops.layer_norm import dropout_add_layer_norm, layer_norm_from_residual
from .mha import MHA
from .mlp import Mlp


class Block(nn.Module):

    def __init__(self, dim, mixer_cls=None, mlp_cls=None, norm_cls=nn.LayerNorm, dropout_cls=nn.Dropout,
                 prenorm=True, resid_dropout=0., drop_path=0., fused_dropout_add_ln=False, return_residual=False,
                 sequence_parallel=False, fuse_residual_add="none"):
        super().__init__()
        if not prenorm:
            raise RuntimeError("post-norm blocks are out of scope (GPT/Backpack use prenorm=True)")
        if drop_path != 0. or return_residual or sequence_parallel:
            raise RuntimeError("drop_path / return_residual / sequence_parallel are training-only features")
        self.prenorm = True
        self.fused_dropout_add_ln = fused_dropout_add_ln
        self.return_residual = False
        # move the residual add into the branch GEMM's epilogue when possible (see the module docstring); the
        # residual tensor handed to forward() is then updated IN PLACE
        # Measured in the Backpack-Small step (A/B in one run, profiles/): 24.9 ms with and without; the GEMM takes
        # over exactly the traffic the LayerNorm sheds (out_proj becomes HBM-bound at 105 us), so the default stays
        # the two-kernel path of the reference.
        if fuse_residual_add not in ("all", "mixer", "mlp", "none"):
            raise ValueError('fuse_residual_add must be "all", "mixer", "mlp" or "none"')
        self.fuse_residual_add = fuse_residual_add   # model factories pass `config.fuse_residual_add`
        if mixer_cls is None:
            mixer_cls = partial(MHA, num_heads=dim // 64)
        if mlp_cls is None:
            mlp_cls = partial(Mlp, hidden_features=4 * dim)
        self.mixer = mixer_cls(dim)
        self.dropout1 = dropout_cls(resid_dropout)
        self.norm1 = norm_cls(dim)
        self.mlp = mlp_cls(dim)
        if not isinstance(self.mlp, nn.Identity):
            self.dropout2 = dropout_cls(resid_dropout)
            self.norm2 = norm_cls(dim)
        if self.fused_dropout_add_ln and not isinstance(self.norm1, nn.LayerNorm):
            raise RuntimeError("fused_dropout_add_ln needs nn.LayerNorm norms")

    def _add_norm(self, branch: Tensor, residual: Tensor, norm: nn.LayerNorm, dropout: nn.Module):
        if self.fused_dropout_add_ln:
            return dropout_add_layer_norm(branch, residual, norm.weight, norm.bias,
                                          dropout.p if self.training else 0.0, norm.eps, prenorm=True)
        residual = dropout(branch) + residual
        return norm(residual.to(dtype=norm.weight.dtype)), residual

    def forward(self, hidden_states: Tensor, residual: Optional[Tensor] = None, mixer_kwargs=None):
        """hidden_states = LayerNorm(residual) on entry; returns the updated (hidden_states, residual)."""
        if residual is None:
            raise RuntimeError("prenorm Block needs the residual stream")
        mixer_kwargs = mixer_kwargs if mixer_kwargs is not None else {}
        fuse = self.fused_dropout_add_ln and not self.training
        fuse_mixer = fuse and self.fuse_residual_add in ("all", "mixer")
        fuse_mlp = fuse and self.fuse_residual_add in ("all", "mlp")
        if (fuse_mixer and isinstance(self.mixer, MHA) and isinstance(self.mixer.out_proj, FusedDense)
                and can_fuse_residual(hidden_states, self.mixer.out_proj.weight, residual)):
            residual = self.mixer(hidden_states, residual_out=residual, **mixer_kwargs)
            hidden_states = layer_norm_from_residual(residual, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        else:
            mixer_out = self.mixer(hidden_states, **mixer_kwargs)
            hidden_states, residual = self._add_norm(mixer_out, residual, self.norm1, self.dropout1)
        if not isinstance(self.mlp, nn.Identity):
            if (fuse_mlp and isinstance(self.mlp, FusedDenseGeluDense)
                    and can_fuse_residual(hidden_states, self.mlp.fc2.weight, residual)):
                residual = self.mlp.forward_into_residual(hidden_states, residual)
                hidden_states = layer_norm_from_residual(residual, self.norm2.weight, self.norm2.bias, self.norm2.eps)
            else:
                mlp_out = self.mlp(hidden_states)
                hidden_states, residual = self._add_norm(mlp_out, residual, self.norm2, self.dropout2)
        return hidden_states, residual
