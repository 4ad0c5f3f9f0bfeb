"""Pre-norm residual `Block` (flash_attn/modules/block.py:22-106).

    mixer_out = mixer(h);  res = mixer_out + res;  h = norm1(res)
    mlp_out   = mlp(h);    res = mlp_out + res;    h = norm2(res)

With `fused_dropout_add_ln` each "add + LayerNorm" pair is one bp_ln_residual_fwd launch that keeps the
residual stream in fp32 (block.py:84-88, 101-105); without it the un-fused PyTorch sequence of the
reference is kept, including its rounding of the residual to the weight dtype before LayerNorm (:76).
Dropout and stochastic depth are identities in eval mode and not modelled; post-norm is out of scope.
"""
from __future__ import annotations

from functools import partial
from typing import Optional

import torch.nn as nn
from torch import Tensor

from ..ops.layer_norm import dropout_add_layer_norm
from .mha import MHA
from .mlp import Mlp


class Block(nn.Module):

    def __init__(self, dim, mixer_cls=None, mlp_cls=None, norm_cls=nn.LayerNorm, dropout_cls=nn.Dropout,
                 prenorm=True, resid_dropout=0., drop_path=0., fused_dropout_add_ln=False, return_residual=False,
                 sequence_parallel=False):
        super().__init__()
        if not prenorm:
            raise RuntimeError("post-norm blocks are out of scope (GPT/Backpack use prenorm=True)")
        if drop_path != 0. or return_residual or sequence_parallel:
            raise RuntimeError("drop_path / return_residual / sequence_parallel are training-only features")
        self.prenorm = True
        self.fused_dropout_add_ln = fused_dropout_add_ln
        self.return_residual = False
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
        mixer_out = self.mixer(hidden_states, **(mixer_kwargs if mixer_kwargs is not None else {}))
        hidden_states, residual = self._add_norm(mixer_out, residual, self.norm1, self.dropout1)
        if not isinstance(self.mlp, nn.Identity):
            mlp_out = self.mlp(hidden_states)
            hidden_states, residual = self._add_norm(mlp_out, residual, self.norm2, self.dropout2)
        return hidden_states, residual
