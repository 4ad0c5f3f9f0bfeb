"""Greedy decoding / sampling with the reference's interface (training/src/utils/generation.py:11-92 for Backpacks,
flash_attn/utils/generation.py:11-65 for the KV-cache machinery), plus the incremental path the reference lacks for
Backpacks.

The reference's Backpack loop re-runs the whole forward over the growing prefix for every generated token
(generation.py:34-44, 62-72) -- O(n^2) trunk work and an O(n^2) sense-mix per token.  Here the prompt is run once
(filling the trunk's KV caches, the contextualisation-key cache and the token-id cache held by `InferenceParams`) and
each further token costs one seqlen-1 step: GEMMs on one row per sequence, `decode_attention` per layer and one
`sense_mix_decode`.  `incremental=False` keeps the reference's re-run loop (with the LM head restricted to the last
position) for models or settings the incremental path does not cover.

A decode step is ~130 small kernels, i.e. launch-bound when driven from Python (2.6 ms per token on a B200 host at any
batch size).  With `cuda_graph=True` (the default for greedy decoding on fused CUDA models) the whole step -- model,
arg-max, the append to the output buffer and the advance of the position counters -- is captured ONCE in a CUDA graph
and replayed per token with no host synchronisation: the write position and the context lengths live in device tensors
(`InferenceParams.cache_position` / `cache_lengths`), the caches are written with index_copy_ and the decode kernels
read the lengths from the device.

Return value: the reference returns `sequences` of max_length - 1 tokens -- its loop drops the last sampled token
(generation.py:38-43: the token sampled in the final iteration is never appended) -- and its greedy loop only works
for batch 1 (`next_token.unsqueeze(0)`, :70).  This module returns all max_length tokens for any batch size, as the
reference's docstring promises; the first max_length - 1 columns are the reference's.
"""
from __future__ import annotations

from collections import namedtuple
from dataclasses import dataclass, field
from typing import Optional

import torch

# transformers 5 no longer ships GreedySearchDecoderOnlyOutput / SampleDecoderOnlyOutput; same two fields
GreedySearchDecoderOnlyOutput = namedtuple("GreedySearchDecoderOnlyOutput", ["sequences", "scores"])
SampleDecoderOnlyOutput = namedtuple("SampleDecoderOnlyOutput", ["sequences", "scores"])


@dataclass
class InferenceParams:
    """Mirrors flash_attn/utils/generation.py:11-19.  `key_value_memory_dict` maps a layer index to that layer's
    (max_batch_size, max_sequence_len, 2, nheads, headdim) KV cache; the Backpack model adds the string keys
    "backpack.ctx_k" (max_batch, max_seq, nv, dk) and "backpack.ids" (max_batch, max_seq).

    Device-side offsets (no counterpart in the reference; both None = the reference's host-side bookkeeping): when
    `cache_position` (int64, shape (1,), on the GPU) is set, a call is a single-token decode step that writes its K/V
    at that position and attends to the first `cache_lengths[b]` (int32, shape (max_batch_size,)) cached positions;
    `sequence_len_offset` is ignored.  Nothing in such a step depends on host integers, so it can be captured in a
    CUDA graph and replayed while the two tensors are advanced on the device."""
    max_sequence_len: int
    max_batch_size: int
    sequence_len_offset: int = 0
    batch_size_offset: int = 0
    key_value_memory_dict: dict = field(default_factory=dict)
    cache_position: Optional[torch.Tensor] = None
    cache_lengths: Optional[torch.Tensor] = None


def _pick(logits, do_sample):
    if do_sample:
        return torch.distributions.Categorical(logits=torch.log_softmax(logits.float(), dim=-1)).sample()
    return torch.argmax(logits, dim=-1)


def _decode_graphed(input_ids, model, max_length, want_scores=True):
    """Greedy decoding with the single-token step captured in a CUDA graph (see the module docstring)."""
    batch_size, seqlen_og = input_ids.shape
    dev = input_ids.device
    n_new = max_length - seqlen_og
    with torch.inference_mode():
        params = InferenceParams(max_sequence_len=max_length, max_batch_size=batch_size)
        logits = model(input_ids, inference_params=params, num_last_tokens=1).logits[:, -1]
        tokens = torch.empty((batch_size, n_new), dtype=torch.long, device=dev)
        scores = torch.empty((n_new, *logits.shape), dtype=logits.dtype, device=dev) if want_scores else None
        first = logits.argmax(dim=-1, keepdim=True)
        tokens[:, :1] = first
        if want_scores:
            scores[0] = logits
        if n_new > 1:
            cur = first.clone()                                                           # (b, 1) token fed to the step
            pos = torch.full((batch_size, 1), seqlen_og, dtype=torch.long, device=dev)    # its position id
            params.cache_position = torch.full((1,), seqlen_og, dtype=torch.long, device=dev)
            params.cache_lengths = torch.full((batch_size,), seqlen_og + 1, dtype=torch.int32, device=dev)
            slot = torch.ones(1, dtype=torch.long, device=dev)                            # column of `tokens` to fill

            def step():
                lg = model(cur, position_ids=pos, inference_params=params, num_last_tokens=1).logits[:, -1]
                nxt = lg.argmax(dim=-1, keepdim=True)
                tokens.index_copy_(1, slot, nxt)
                if want_scores:
                    scores.index_copy_(0, slot, lg.unsqueeze(0))
                cur.copy_(nxt)
                pos.add_(1)
                params.cache_position.add_(1)
                params.cache_lengths.add_(1)
                slot.add_(1)

            # one eager step on a side stream (library handles, workspaces, the sense table), then rewind the counters:
            # what it wrote (cache row `seqlen_og`, tokens[:, 1]) is exactly what the first replay writes again
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream(dev).wait_stream(side)
            cur.copy_(first)
            pos.fill_(seqlen_og)
            params.cache_position.fill_(seqlen_og)
            params.cache_lengths.fill_(seqlen_og + 1)
            slot.fill_(1)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            for _ in range(n_new - 1):
                graph.replay()
        sequences = torch.cat([input_ids, tokens], dim=1)
        return sequences, (tuple(scores.unbind(0)) if want_scores else None)


def _decode(input_ids, model, max_length, do_sample, incremental, cuda_graph=None, want_scores=True):
    batch_size, seqlen_og = input_ids.shape
    if max_length <= seqlen_og:
        raise ValueError(f"max_length ({max_length}) must exceed the prompt length ({seqlen_og})")
    if cuda_graph is None:
        cuda_graph = (incremental and not do_sample and input_ids.is_cuda
                      and bool(getattr(model, "graphed_decode_ok", lambda: False)()))
    if cuda_graph:
        if not incremental or do_sample:
            raise ValueError("cuda_graph=True needs incremental greedy decoding")
        return _decode_graphed(input_ids, model, max_length, want_scores)
    scores, new_tokens = [], []
    with torch.inference_mode():
        if incremental:
            params = InferenceParams(max_sequence_len=max_length, max_batch_size=batch_size)
            logits = model(input_ids, inference_params=params, num_last_tokens=1).logits[:, -1]
            params.sequence_len_offset = seqlen_og
            while True:
                scores.append(logits)
                next_token = _pick(logits, do_sample)
                new_tokens.append(next_token)
                if seqlen_og + len(new_tokens) >= max_length:
                    break
                position_ids = torch.full((batch_size, 1), params.sequence_len_offset, dtype=torch.long,
                                          device=input_ids.device)
                logits = model(next_token.unsqueeze(1), position_ids=position_ids, inference_params=params,
                               num_last_tokens=1).logits[:, -1]
                params.sequence_len_offset += 1
        else:
            ids = input_ids
            while True:
                logits = model(ids, num_last_tokens=1).logits[:, -1]
                scores.append(logits)
                next_token = _pick(logits, do_sample)
                new_tokens.append(next_token)
                if seqlen_og + len(new_tokens) >= max_length:
                    break
                ids = torch.cat((ids, next_token.unsqueeze(1)), dim=1)
    return torch.cat([input_ids, torch.stack(new_tokens, dim=1)], dim=1), tuple(scores)


def greedy_decode(input_ids, model, max_length, incremental=True, cuda_graph=None, output_scores=True):
    """input_ids: (batch, seq_len), every sequence of the same length.  Returns `sequences` (batch, max_length) and
    `scores`, a tuple of (batch, vocab_size) logits, one per generated token.  `cuda_graph`: None = use a captured
    decode step when the model supports it (`model.graphed_decode_ok()`), True / False to force."""
    return GreedySearchDecoderOnlyOutput(*_decode(input_ids, model, max_length, False, incremental, cuda_graph,
                                                  output_scores))


def sample(input_ids, model, max_length, incremental=True):
    """Ancestral sampling from softmax(logits) (generation.py:22-48); uses torch's global RNG like the reference."""
    return SampleDecoderOnlyOutput(*_decode(input_ids, model, max_length, True, incremental))


class GenerationMixin:
    """`generate` / `sample` with the reference's signature (generation.py:80-92); `incremental` selects the KV-cache
    path (default) or the reference's prefix re-run."""

    def generate(self, input_ids, max_length, return_dict_in_generate=False, output_scores=False, incremental=True,
                 cuda_graph=None):
        output = greedy_decode(input_ids, self, max_length, incremental=incremental, cuda_graph=cuda_graph,
                               output_scores=output_scores)
        if not output_scores:
            output = output._replace(scores=None)
        return output if return_dict_in_generate else output.sequences

    def sample(self, input_ids, max_length, return_dict_in_generate=False, output_scores=False, incremental=True):
        output = sample(input_ids, self, max_length, incremental=incremental)
        if not output_scores:
            output = output._replace(scores=None)
        return output if return_dict_in_generate else output.sequences
