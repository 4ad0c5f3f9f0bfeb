"""Replay the forward pass as ONE CUDA graph.

A Backpack-Small forward is ~110 kernel launches (95 of this library, the LM-head GEMM, gathers and a few element-wise
ATen kernels); every shape is static for a fixed (batch, seqlen), the library never allocates and never synchronises,
and its only launch-time state -- the attention scheduler's ticket counter -- is PER LAUNCH: a slot of a device-side
ring zeroed by a memset node in front of the kernel (csrc/bp_fmha_fwd.cu), so the whole launch sequence can be captured
once and replayed, several graphs can replay concurrently on different streams next to eager launches, and a capture
that would exhaust the slot pool fails with an error instead of sharing a counter.  Replaying removes the gaps between
dependent launches (measured on B200: 25.5 -> 25.0 ms per step).  The decode step of utils/generation.py is captured
the same way.  The reference has no counterpart (its generation loop re-launches everything,
training/src/utils/generation.py:34-44); this is an opt-in serving helper, not a different compute path -- the
same kernels run on the same buffers.

    fwd = GraphedForward(model, example_ids)      # captures model(example_ids)
    logits = fwd(ids)                             # copies ids into the static input, replays, returns static logits
"""
from __future__ import annotations

import torch


class GraphedForward:

    def __init__(self, model: torch.nn.Module, example_ids: torch.Tensor, warmup: int = 2):
        if not example_ids.is_cuda:
            raise RuntimeError("GraphedForward needs CUDA inputs (there is no CPU path)")
        if torch.is_grad_enabled():
            raise RuntimeError("capture the forward under torch.inference_mode() / no_grad()")
        self.model = model
        self.static_ids = example_ids.clone()
        side = torch.cuda.Stream(device=example_ids.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):              # warm-up off the capturing stream (lazy initialisations)
            for _ in range(warmup):
                model(self.static_ids)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.logits = model(self.static_ids).logits

    def __call__(self, input_ids: torch.Tensor | None = None) -> torch.Tensor:
        """input_ids None: replay on whatever the static input holds.  The returned tensor is the graph's static
        output buffer: it is overwritten by the next call."""
        if input_ids is not None:
            if input_ids.shape != self.static_ids.shape:
                raise RuntimeError(f"graph captured for ids of shape {tuple(self.static_ids.shape)}, got "
                                   f"{tuple(input_ids.shape)}")
            self.static_ids.copy_(input_ids, non_blocking=True)
        self.graph.replay()
        return self.logits
