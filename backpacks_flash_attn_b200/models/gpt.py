"""GPT trunk with the reference's interface (flash_attn/models/gpt.py:44-282): `create_mixer_cls`,
`create_mlp_cls`, `create_block`, `GPTModel`, `GPTLMHeadModel`.

Config attributes are read exactly as the reference reads them (`use_flash_attn`, `fused_bias_fc`,
`fused_dense_gelu_dense`, `fused_dropout_add_ln`, `pad_vocab_size_multiple`,
`scale_attn_by_inverse_layer_idx`, `rotary_emb_fraction`, ...) and state-dict keys are unchanged, so a
checkpoint trained with the reference loads into these modules.  Tensor parallelism (process_group) is out
of scope: the path shards the batch only.
"""
from __future__ import annotations

import math
from collections import namedtuple
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F
from transformers import GPT2Config

from ..modules.block import Block
from ..modules.embedding import GPT2Embeddings
from ..modules.mha import MHA
from ..modules.mlp import Mlp
from ..ops.fused_dense import FusedDenseGeluDense
from ..ops.layer_norm import dropout_add_layer_norm
from ..utils.generation import GenerationMixin

CausalLMOutput = namedtuple("CausalLMOutput", ["logits"])


def _no_tp(process_group):
    if process_group is not None:
        raise RuntimeError("tensor / sequence parallelism is out of scope; shard the batch across GPUs instead")


def create_mixer_cls(config, layer_idx=None, process_group=None, device=None, dtype=None):
    """softmax scale = head_dim^-0.5, divided by (layer_idx + 1) when scale_attn_by_inverse_layer_idx
    (gpt.py:46-50)."""
    _no_tp(process_group)
    factory_kwargs = {"device": device, "dtype": dtype}
    head_dim = getattr(config, "head_dim", config.hidden_size // config.num_attention_heads)
    softmax_scale = 1.0 if not config.scale_attn_weights else head_dim ** (-0.5)
    if config.scale_attn_by_inverse_layer_idx:
        if layer_idx is None:
            raise RuntimeError("scale_attn_by_inverse_layer_idx needs layer_idx")
        softmax_scale /= float(layer_idx + 1)
    if getattr(config, "attn_dwconv", False):
        raise RuntimeError("attn_dwconv is out of scope")
    rotary_emb_dim = int(getattr(config, "rotary_emb_fraction", 0.0) * head_dim)
    return partial(MHA, num_heads=config.num_attention_heads, dropout=config.attn_pdrop,
                   softmax_scale=softmax_scale, causal=True, layer_idx=layer_idx, rotary_emb_dim=rotary_emb_dim,
                   rotary_emb_scale_base=getattr(config, "rotary_emb_scale_base", 0),
                   use_flash_attn=getattr(config, "use_flash_attn", False),
                   fused_bias_fc=getattr(config, "fused_bias_fc", False), **factory_kwargs)


def _activation(config):
    if config.activation_function == "sqrelu":
        raise RuntimeError("sqrelu (Triton fused MLP) is out of scope")
    approximate = "tanh" if config.activation_function in ["gelu_new", "gelu_fast"] else "none"
    return partial(F.gelu, approximate=approximate)


def create_mlp_cls(config, layer_idx=None, process_group=None, device=None, dtype=None,
                   inner_dim=None, out_features=None):
    """Mlp or FusedDenseGeluDense according to `fused_dense_gelu_dense` (gpt.py:72-108)."""
    _no_tp(process_group)
    factory_kwargs = {"device": device, "dtype": dtype}
    if inner_dim is None:
        inner_dim = config.n_inner if config.n_inner is not None else 4 * config.hidden_size
    if getattr(config, "fused_dense_sqrelu_dense", False):
        raise RuntimeError("fused_dense_sqrelu_dense (Triton) is out of scope")
    if getattr(config, "fused_dense_gelu_dense", False):
        if config.activation_function not in ["gelu_new", "gelu_fast"]:
            raise RuntimeError("fused_dense_gelu_dense only supports approximate gelu")
        return partial(FusedDenseGeluDense, hidden_features=inner_dim, out_features=out_features, **factory_kwargs)
    return partial(Mlp, hidden_features=inner_dim, out_features=out_features, activation=_activation(config),
                   **factory_kwargs)


def create_block(config, layer_idx=None, process_group=None, device=None, dtype=None):
    _no_tp(process_group)
    factory_kwargs = {"device": device, "dtype": dtype}
    mixer_cls = create_mixer_cls(config, layer_idx, **factory_kwargs)
    mlp_cls = create_mlp_cls(config, layer_idx, **factory_kwargs)
    norm_cls = partial(nn.LayerNorm, eps=config.layer_norm_epsilon, **factory_kwargs)
    block = Block(config.hidden_size, mixer_cls, mlp_cls, norm_cls=norm_cls, prenorm=True,
                  resid_dropout=config.resid_pdrop,
                  fused_dropout_add_ln=getattr(config, "fused_dropout_add_ln", False),
                  fuse_residual_add=getattr(config, "fuse_residual_add", "none"))
    block.layer_idx = layer_idx
    return block


def _init_weights(module, n_layer, initializer_range=0.02, rescale_prenorm_residual=True):
    """GPT-2 initialisation (gpt.py:154-172): N(0, 0.02), residual projections scaled by 1/sqrt(2 n_layer)."""
    if isinstance(module, nn.Linear):
        nn.init.normal_(module.weight, std=initializer_range)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.Embedding):
        nn.init.normal_(module.weight, std=initializer_range)
    if rescale_prenorm_residual:
        for name, p in module.named_parameters():
            if name in ["out_proj.weight", "fc2.weight"]:
                nn.init.normal_(p, mean=0.0, std=initializer_range / math.sqrt(2 * n_layer))


def pad_vocab(config):
    """Round config.vocab_size up in place (gpt.py:182-185)."""
    mult = getattr(config, "pad_vocab_size_multiple", 1)
    if config.vocab_size % mult != 0:
        config.vocab_size += mult - (config.vocab_size % mult)
    return mult


def first_layer_norm(hidden_states, ln, dropout, fused, training):
    """residual = fp32(embeddings); hidden = ln_0(residual)  (gpt.py:232-240)."""
    if not fused:
        residual = dropout(hidden_states).float()
        return ln(residual.to(dtype=ln.weight.dtype)), residual
    return dropout_add_layer_norm(hidden_states, None, ln.weight, ln.bias, dropout.p if training else 0.0, ln.eps,
                                  prenorm=True, residual_in_fp32=True)


class GPTPreTrainedModel(nn.Module):

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        if not isinstance(config, GPT2Config):
            raise ValueError(f"Parameter config in `{self.__class__.__name__}(config)` should be an instance of "
                             "class `GPT2Config`.")
        self.config = config


class GPTModel(GPTPreTrainedModel):
    """Embeddings -> ln_0 -> n_layer prenorm Blocks; returns the last block's norm2 output (the HF ln_f)."""

    def __init__(self, config: GPT2Config, process_group=None, device=None, dtype=None):
        super().__init__(config)
        _no_tp(process_group)
        factory_kwargs = {"device": device, "dtype": dtype}
        self.process_group = None
        if config.activation_function not in ["gelu", "gelu_new", "gelu_fast"]:
            raise RuntimeError(f"unsupported activation {config.activation_function}")
        self.pad_vocab_size_multiple = pad_vocab(config)
        self.embeddings = GPT2Embeddings(config.hidden_size, config.vocab_size, config.max_position_embeddings,
                                         **factory_kwargs)
        self.emb_drop = nn.Dropout(config.embd_pdrop)
        self.fused_dropout_add_ln = getattr(config, "fused_dropout_add_ln", False)
        self.ln_0 = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_epsilon, **factory_kwargs)
        self.layers = nn.ModuleList([create_block(config, layer_idx=i, **factory_kwargs)
                                     for i in range(config.num_hidden_layers)])
        self.apply(partial(_init_weights, n_layer=config.num_hidden_layers,
                           initializer_range=config.initializer_range))

    def forward(self, input_ids, position_ids=None, inference_params=None):
        hidden_states = self.embeddings(input_ids, position_ids=position_ids)
        hidden_states, residual = first_layer_norm(hidden_states, self.ln_0, self.emb_drop,
                                                   self.fused_dropout_add_ln, self.training)
        # gpt.py:241-245: the KV caches travel to the mixers as a keyword argument
        mixer_kwargs = {"inference_params": inference_params} if inference_params is not None else None
        for layer in self.layers:
            hidden_states, residual = layer(hidden_states, residual, mixer_kwargs=mixer_kwargs)
        return hidden_states


class GPTLMHeadModel(GPTPreTrainedModel, GenerationMixin):

    def __init__(self, config: GPT2Config, process_group=None, device=None, dtype=None):
        super().__init__(config)
        _no_tp(process_group)
        factory_kwargs = {"device": device, "dtype": dtype}
        self.transformer = GPTModel(config, **factory_kwargs)
        self.lm_head = nn.Linear(config.n_embd, config.vocab_size, bias=False, **factory_kwargs)
        self.apply(partial(_init_weights, n_layer=config.num_hidden_layers,
                           initializer_range=config.initializer_range))
        self.tie_weights()

    def tie_weights(self):
        self.lm_head.weight = self.transformer.embeddings.word_embeddings.weight

    def graphed_decode_ok(self):
        """Whether a decode step can run with device-side offsets, i.e. inside a CUDA graph (utils/generation.py)."""
        cfg = self.config
        return bool(getattr(cfg, "use_flash_attn", False) and not self.training
                    and getattr(cfg, "rotary_emb_fraction", 0.0) == 0.0 and cfg.n_embd // cfg.n_head in (64, 128))

    def forward(self, input_ids, position_ids=None, inference_params=None, num_last_tokens=0):
        """inference_params: KV caches for generation (gpt.py:273-281).  num_last_tokens > 0 projects only the last
        positions to the vocabulary (the generation loop reads logits[:, -1] only)."""
        hidden_states = self.transformer(input_ids, position_ids=position_ids, inference_params=inference_params)
        if num_last_tokens > 0:
            hidden_states = hidden_states[:, -num_last_tokens:]
        if getattr(self.config, "fused_bias_fc", False):
            from ..ops.fused_dense import linear
            return CausalLMOutput(logits=linear(hidden_states, self.lm_head.weight, self.lm_head.bias))
        return CausalLMOutput(logits=self.lm_head(hidden_states))
