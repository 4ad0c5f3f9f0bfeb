"""Greedy decoding / sampling with the reference's interface (training/src/utils/generation.py:11-92 for Backpacks,
flash_attn/utils/generation.py:11-65 for the KV-cache machinery), plus the incremental path the reference lacks for
Backpacks.

The reference's Backpack loop re-runs the whole forward over the growing prefix for every generated token
(generation.py:34-44, 62-72) -- O(n^2) trunk work and an O(n^2) sense-mix per token.  Here the prompt is run once
(filling the trunk's KV caches, the contextualisation-key cache and the token-id cache held by `InferenceParams`) and
each further token costs one seqlen-1 step: GEMMs on one row per sequence, `decode_attention` per layer and one
`sense_mix_decode`.  `incremental=False` keeps the reference's re-run loop (with the LM head restricted to the last
position) for models or settings the incremental path does not cover.

Return value: the reference returns `sequences` of max_length - 1 tokens -- its loop drops the last sampled token
(generation.py:38-43: the token sampled in the final iteration is never appended) -- and its greedy loop only works
for batch 1 (`next_token.unsqueeze(0)`, :70).  This module returns all max_length tokens for any batch size, as the
reference's docstring promises; the first max_length - 1 columns are the reference's.
"""
from __future__ import annotations

from collections import namedtuple
from dataclasses import dataclass, field

import torch

# transformers 5 no longer ships GreedySearchDecoderOnlyOutput / SampleDecoderOnlyOutput; same two fields
GreedySearchDecoderOnlyOutput = namedtuple("GreedySearchDecoderOnlyOutput", ["sequences", "scores"])
SampleDecoderOnlyOutput = namedtuple("SampleDecoderOnlyOutput", ["sequences", "scores"])


@dataclass
class InferenceParams:
    """Mirrors flash_attn/utils/generation.py:11-19.  `key_value_memory_dict` maps a layer index to that layer's
    (max_batch_size, max_sequence_len, 2, nheads, headdim) KV cache; the Backpack model adds the string keys
    "backpack.ctx_k" (max_batch, max_seq, nv, dk) and "backpack.ids" (max_batch, max_seq)."""
    max_sequence_len: int
    max_batch_size: int
    sequence_len_offset: int = 0
    batch_size_offset: int = 0
    key_value_memory_dict: dict = field(default_factory=dict)


def _pick(logits, do_sample):
    if do_sample:
        return torch.distributions.Categorical(logits=torch.log_softmax(logits.float(), dim=-1)).sample()
    return torch.argmax(logits, dim=-1)


def _decode(input_ids, model, max_length, do_sample, incremental):
    batch_size, seqlen_og = input_ids.shape
    if max_length <= seqlen_og:
        raise ValueError(f"max_length ({max_length}) must exceed the prompt length ({seqlen_og})")
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


def greedy_decode(input_ids, model, max_length, incremental=True):
    """input_ids: (batch, seq_len), every sequence of the same length.  Returns `sequences` (batch, max_length) and
    `scores`, a tuple of (batch, vocab_size) logits, one per generated token."""
    return GreedySearchDecoderOnlyOutput(*_decode(input_ids, model, max_length, False, incremental))


def sample(input_ids, model, max_length, incremental=True):
    """Ancestral sampling from softmax(logits) (generation.py:22-48); uses torch's global RNG like the reference."""
    return SampleDecoderOnlyOutput(*_decode(input_ids, model, max_length, True, incremental))


class GenerationMixin:
    """`generate` / `sample` with the reference's signature (generation.py:80-92); `incremental` selects the KV-cache
    path (default) or the reference's prefix re-run."""

    def generate(self, input_ids, max_length, return_dict_in_generate=False, output_scores=False, incremental=True):
        output = greedy_decode(input_ids, self, max_length, incremental=incremental)
        if not output_scores:
            output = output._replace(scores=None)
        return output if return_dict_in_generate else output.sequences

    def sample(self, input_ids, max_length, return_dict_in_generate=False, output_scores=False, incremental=True):
        output = sample(input_ids, self, max_length, incremental=incremental)
        if not output_scores:
            output = output._replace(scores=None)
        return output if return_dict_in_generate else output.sequences
