"""Synthetic, RNG-order-independent weights for benchmarks and parity tests (there is no network for
checkpoints).  Recipe of SURVEY.md §8c: every parameter is drawn from a generator seeded with crc32 of its
canonical name; matrices are fan-in scaled, LayerNorm gains are 1 + 0.1 N(0,1), the rest 0.02 N(0,1)."""
from __future__ import annotations

import zlib

import torch


def name_seeded_(model: torch.nn.Module) -> torch.nn.Module:
    """Fill `model`'s parameters in place (tied parameters are visited once, under their first name)."""
    with torch.no_grad():
        for name, p in sorted(dict(model.named_parameters()).items()):
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
            w = torch.randn(p.shape, generator=g)
            if p.dim() == 2:
                w = w * p.shape[1] ** -0.5
            elif name.endswith(("ln_0.weight", "norm1.weight", "norm2.weight")):
                w = 1 + 0.1 * w
            else:
                w = 0.02 * w
            p.copy_(w.to(device=p.device, dtype=p.dtype))
    return model


def parameter_checksum(model: torch.nn.Module) -> torch.Tensor:
    """Order-independent fp64 checksum of all parameters (used to prove data-parallel replicas match)."""
    total = torch.zeros((), dtype=torch.float64, device=next(model.parameters()).device)
    for _, p in sorted(dict(model.named_parameters()).items()):
        total += p.detach().double().abs().sum()
    return total
