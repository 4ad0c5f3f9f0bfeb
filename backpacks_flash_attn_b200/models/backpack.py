"""Backpack language model with the reference's interface (training/src/models/backpack.py):
`BackpackConfig` (:146-154), `ContextSelfAttn` (:94-122), `BackpackContentModule` (:207-276),
`BackpackModel` (:278-314), `BackpackLMHeadModel` (:318-351).

What differs from the reference is only HOW `BackpackModel.forward` evaluates
`torch.sum(contextualization @ content, dim=1)` (:313): with `use_flash_attn` (or the explicit
`fused_sense_mix` config attribute) the (b, nv, s, s) weights are never materialised -- the projected
(q, k) of `ContextSelfAttn` and the content tensor go straight into the fused sense-mix kernels.  The
sub-module API the analysis scripts use (training/src/models/intervened_models.py:77-101,
training/src/run_simlex.py:179-184) is preserved: `transformer.gpt2_model(ids)`,
`transformer.contextualization_attn(h) -> (b, nv, s, s)` (eager), `transformer.content_model(ids) ->
(b, nv, s, d)`, and `transformer.sense_mix(h, content)` is offered for edited content tensors.
State-dict keys are identical to the reference's.
"""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn as nn
from transformers import GPT2Config

from ..modules.block import Block
from ..ops.fused_dense import FusedDense, linear
from ..ops.decode import sense_mix_decode
from ..ops.sense_mix import sense_mix, sense_mix_table
from ..utils.generation import GenerationMixin
from .gpt import (CausalLMOutput, GPTModel, GPTPreTrainedModel, _init_weights, _no_tp, create_mlp_cls,
                  first_layer_norm, pad_vocab)


class BackpackConfig(GPT2Config):

    def __init__(self, num_content_vectors=16, **kwargs):
        self.num_content_vectors = num_content_vectors
        super().__init__(**kwargs)


def create_content_mlp_cls(config, layer_idx=None, expand_out=False, process_group=None, device=None, dtype=None):
    """MLP factory of the content model (backpack.py:53-92): inner = n_inner or 4d, or d when
    `shrink_final_inner`; out = nv*d when expand_out else d."""
    _no_tp(process_group)
    inner_dim = config.n_inner if config.n_inner is not None else 4 * config.hidden_size
    inner_dim = config.hidden_size if getattr(config, "shrink_final_inner", None) else inner_dim
    outer_dim = config.num_content_vectors * config.hidden_size if expand_out else config.hidden_size
    return create_mlp_cls(config, layer_idx, device=device, dtype=dtype, inner_dim=inner_dim, out_features=outer_dim)


class ContextSelfAttn(nn.Module):
    """num_content_vectors causal attention maps per token pair (backpack.py:94-122).
    Parameter: Wqkv (embed_dim -> 2*embed_dim), q = first half, k = second half, split into nv senses."""

    def __init__(self, num_content_vectors, embed_dim, device=None, dtype=None):
        super().__init__()
        self.Wqkv = FusedDense(embed_dim, 2 * embed_dim, device=device, dtype=dtype)
        self.num_content_vectors = num_content_vectors
        self.softmax_scale = None

    def project_qk(self, encoded):
        """(b, s, d) -> (b, s, 2, nv, d // nv)"""
        b, s, d = encoded.shape
        return self.Wqkv(encoded).reshape(b, s, 2, self.num_content_vectors, d // self.num_content_vectors)

    def forward(self, encoded):
        """Eager contextualisation weights alpha (b, nv, s, s), exactly the reference's composition."""
        qk = self.project_qk(encoded)
        s = qk.shape[1]
        q, k = qk.unbind(dim=2)
        softmax_scale = self.softmax_scale or 1.0 / math.sqrt(q.shape[-1])
        scores = torch.einsum("bthd,bshd->bhts", q, k * softmax_scale)
        causal_mask = torch.triu(torch.full((s, s), -10000.0, device=scores.device), 1)
        scores = scores + causal_mask.to(dtype=scores.dtype)
        return torch.softmax(scores, dim=-1, dtype=q.dtype)


class Identity(nn.Identity):

    def forward(self, x, **kwargs):
        return x


def create_nomix_block(config, expand_out=False, layer_idx=None, process_group=None, device=None, dtype=None):
    """A Block whose mixer is the identity (backpack.py:130-143): residual = LN0(e) + e, then the MLP half."""
    _no_tp(process_group)
    factory_kwargs = {"device": device, "dtype": dtype}
    mlp_cls = create_content_mlp_cls(config, layer_idx, expand_out, **factory_kwargs)
    norm_cls = partial(nn.LayerNorm, eps=config.layer_norm_epsilon, **factory_kwargs)
    block = Block(config.hidden_size, Identity, mlp_cls, norm_cls=norm_cls, prenorm=True,
                  resid_dropout=config.resid_pdrop, fused_dropout_add_ln=getattr(config, "fused_dropout_add_ln", False),
                  fuse_residual_add=getattr(config, "fuse_residual_add", "none"))
    block.layer_idx = layer_idx
    return block


class BackpackPreTrainedModel(nn.Module):

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        if not isinstance(config, BackpackConfig):
            raise ValueError(f"Parameter config in `{self.__class__.__name__}(config)` should be a BackpackConfig")
        self.config = config


class BackpackContentModule(nn.Module):
    """Context-free sense vectors C(x) (backpack.py:207-276): word embedding (no positions) -> ln_0 ->
    one Identity-mixer block -> final MLP to nv*d, returned as the (b, nv, s, d) transposed view."""

    def __init__(self, config, num_content_vectors, embeddings, process_group=None, device=None, dtype=None):
        super().__init__()
        _no_tp(process_group)
        factory_kwargs = {"device": device, "dtype": dtype}
        self.num_content_vectors = num_content_vectors
        self.embeddings = embeddings
        self.process_group = None
        self.n_embd = config.n_embd
        self.fused_dropout_add_ln = getattr(config, "fused_dropout_add_ln", False)
        self.ln_0 = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_epsilon, **factory_kwargs)
        n_layers = 1
        self.layers = nn.ModuleList([create_nomix_block(config, layer_idx=i, expand_out=False, **factory_kwargs)
                                     for i in range(n_layers)])
        self.final_mlp = create_content_mlp_cls(config, layer_idx=n_layers + 1, expand_out=True,
                                                **factory_kwargs)(config.n_embd)
        self.emb_drop = nn.Dropout(config.embd_pdrop)
        self.apply(partial(_init_weights, n_layer=n_layers, initializer_range=config.initializer_range))

    def forward(self, input_ids, position_ids=None, inference_params=None):
        hidden_states = self.embeddings.word_embeddings(input_ids)  # no positions (backpack.py:258)
        hidden_states, residual = first_layer_norm(hidden_states, self.ln_0, self.emb_drop,
                                                   self.fused_dropout_add_ln, self.training)
        for layer in self.layers:
            hidden_states, residual = layer(hidden_states, residual)
        hidden_states = self.final_mlp(hidden_states)  # (b, s, nv*d)
        b, s, _ = hidden_states.shape
        return hidden_states.reshape(b, s, self.num_content_vectors, self.n_embd).transpose(1, 2)


class BackpackModel(GPTPreTrainedModel):

    def __init__(self, config: BackpackConfig, process_group=None, device=None, dtype=None):
        super().__init__(config)
        _no_tp(process_group)
        factory_kwargs = {"device": device, "dtype": dtype}
        self.process_group = None
        self.pad_vocab_size_multiple = pad_vocab(config)
        self.num_content_vectors = config.num_content_vectors
        self.gpt2_model = GPTModel(config, **factory_kwargs)
        self.content_model = BackpackContentModule(config, self.num_content_vectors, self.gpt2_model.embeddings,
                                                   **factory_kwargs)
        self.embeddings = self.gpt2_model.embeddings  # shared with the contextualisation trunk
        self.contextualization_attn = ContextSelfAttn(self.num_content_vectors, config.n_embd, **factory_kwargs)
        # fused sense-mix follows use_flash_attn unless the config says otherwise
        self.fused_sense_mix = getattr(config, "fused_sense_mix", getattr(config, "use_flash_attn", False))
        # `config.use_sense_table`: in eval mode serve the sense vectors from a precomputed (vocab, nv, d) table and
        # let the sense-mix kernel gather them by token id; see build_sense_table()
        self.use_sense_table = bool(getattr(config, "use_sense_table", False))
        self.register_buffer("sense_table", None, persistent=False)   # follows .to(); never in the state dict
        self._sense_table_key = None

    # ---- sense-vector table -------------------------------------------------------------------------------------
    def _content_params_key(self):
        """Identity + in-place version of every tensor the content model reads.  load_state_dict, optimizer steps and
        the intervention scripts' weight edits all bump `_version`; .to() replaces the tensors."""
        def version(p):
            try:
                return p._version
            except RuntimeError:      # inference tensors do not track versions
                return -1
        return tuple((p.data_ptr(), version(p), p.dtype, p.device) for p in self.content_model.parameters())

    @torch.no_grad()
    def build_sense_table(self, chunk: int = 8192):
        """Precompute C(x) for every vocabulary item.

        The content model is context-free: it sees the word embedding only (no positions, identity mixer;
        training/src/models/backpack.py:258, :125-143), so for inference `content_model(ids)` is a row gather
        from a (vocab, nv, d) table computed once with the very same kernels.  The analysis scripts of the
        reference rely on the same fact (training/src/run_simlex.py:179-184).  The table is a non-persistent
        buffer (it follows `.to()` and is not part of the state dict) and is rebuilt automatically when a
        content-model parameter has changed since it was built (`load_state_dict`, in-place edits)."""
        emb = self.embeddings.word_embeddings
        vocab, d, nv = emb.num_embeddings, self.config.n_embd, self.num_content_vectors
        table = torch.empty((vocab, nv, d), dtype=emb.weight.dtype, device=emb.weight.device)
        was_training = self.content_model.training
        self.content_model.eval()
        for start in range(0, vocab, chunk):
            ids = torch.arange(start, min(vocab, start + chunk), device=emb.weight.device).unsqueeze(0)
            table[start:start + ids.shape[1]] = self.content_model(ids)[0].transpose(0, 1)
        self.content_model.train(was_training)
        self.sense_table = table
        self._sense_table_key = self._content_params_key()
        return table

    def drop_sense_table(self):
        self.sense_table = None
        self._sense_table_key = None

    def current_sense_table(self, force=False):
        """The table if it is enabled and usable now (eval mode), (re)built when missing or stale; else None.
        `force`: build it even if the config did not ask for it (decode steps need it)."""
        if self.training:
            if force:
                raise RuntimeError("incremental decoding needs eval mode (the sense-vector table is inference-only)")
            return None
        if not (force or self.use_sense_table or self.sense_table is not None):
            return None
        emb = self.embeddings.word_embeddings.weight
        if (self.sense_table is None or self._sense_table_key != self._content_params_key()
                or self.sense_table.device != emb.device or self.sense_table.dtype != emb.dtype):
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the sense-vector table is missing or stale: call build_sense_table() before "
                                   "capturing a CUDA graph")
            self.build_sense_table()
        return self.sense_table

    def content(self, input_ids, position_ids=None, inference_params=None):
        """Sense vectors (b, nv, s, d): the content model, or (eval mode, table enabled) a gather from the table.
        `forward` does not call this in table mode -- the kernel gathers by itself; this is for callers that want
        the tensor (analysis scripts)."""
        table = self.current_sense_table()
        if table is None:
            return self.content_model(input_ids, position_ids, inference_params)
        b, s = input_ids.shape
        rows = torch.nn.functional.embedding(input_ids, table.view(table.shape[0], -1))   # (b, s, nv*d)
        return rows.view(b, s, self.num_content_vectors, self.config.n_embd).transpose(1, 2)

    def sense_mix(self, contextl_hidden_states, content):
        """sum_l alpha_l(contextl_hidden_states) @ content_l without materialising alpha; `content` may be any
        (b, nv, s, d) tensor (e.g. edited sense vectors)."""
        qk = self.contextualization_attn.project_qk(contextl_hidden_states)
        return sense_mix(qk, content, softmax_scale=self.contextualization_attn.softmax_scale)

    def forward(self, input_ids, position_ids=None, inference_params=None):
        contextl_hidden_states = self.gpt2_model(input_ids, position_ids=position_ids,
                                                 inference_params=inference_params)
        if inference_params is not None:
            return self._forward_cached(contextl_hidden_states, input_ids, inference_params)
        if self.fused_sense_mix:
            table = self.current_sense_table()
            if table is not None:
                # C_l(x_j) rows are gathered from the table inside the kernel: no (b, s, nv, d) tensor in HBM
                qk = self.contextualization_attn.project_qk(contextl_hidden_states)
                return sense_mix_table(qk, table, input_ids, softmax_scale=self.contextualization_attn.softmax_scale)
            content = self.content_model(input_ids, position_ids, inference_params)  # (b, nv, s, d)
            return self.sense_mix(contextl_hidden_states, content)
        content = self.content(input_ids, position_ids, inference_params)  # (b, nv, s, d)
        contextualization = self.contextualization_attn(contextl_hidden_states)  # (b, nv, s, s)
        return torch.sum(contextualization @ content, dim=1)

    # ---- incremental decoding (no counterpart in the reference) ----------------------------
    def _decode_caches(self, qk, input_ids, inference_params):
        """Write this call's contextualisation keys and token ids at `sequence_len_offset` into the caches kept in
        `inference_params.key_value_memory_dict` and return the batch slice of both."""
        kvd = inference_params.key_value_memory_dict
        nv, dk = qk.shape[3], qk.shape[4]
        if "backpack.ctx_k" not in kvd:
            kvd["backpack.ctx_k"] = torch.empty(inference_params.max_batch_size, inference_params.max_sequence_len,
                                                nv, dk, dtype=qk.dtype, device=qk.device)
            kvd["backpack.ids"] = torch.zeros(inference_params.max_batch_size, inference_params.max_sequence_len,
                                              dtype=torch.int64, device=qk.device)
        k_cache, ids_cache = kvd["backpack.ctx_k"], kvd["backpack.ids"]
        b0 = inference_params.batch_size_offset
        b1 = b0 + qk.shape[0]
        s0 = inference_params.sequence_len_offset
        s1 = s0 + qk.shape[1]
        if b1 > k_cache.shape[0] or s1 > k_cache.shape[1]:
            raise RuntimeError(f"sense-mix cache of shape {tuple(k_cache.shape)} is too small for batch rows "
                               f"{b0}:{b1}, positions {s0}:{s1}")
        k_cache[b0:b1, s0:s1] = qk[:, :, 1]
        ids_cache[b0:b1, s0:s1] = input_ids
        return k_cache[b0:b1], ids_cache[b0:b1]

    def _forward_cached(self, contextl_hidden_states, input_ids, inference_params):
        """Prompt pass (offset 0): the ordinary forward, with the contextualisation keys and the token ids stored.
        Decode step (one new position): o_n = sum_l sum_{j<=n} alpha_l[n, j] C_l(x_j) from the caches -- the sense
        vectors of the context come from the table (fused) or the content model (eager)."""
        offset = inference_params.sequence_len_offset
        attn = self.contextualization_attn
        qk = attn.project_qk(contextl_hidden_states)               # (b, s, 2, nv, dk)
        dk = qk.shape[-1]
        scale = attn.softmax_scale or 1.0 / math.sqrt(dk)
        if self.fused_sense_mix and dk % 8 != 0:
            qk = torch.nn.functional.pad(qk, (0, (-dk) % 8))        # zero columns leave q.k unchanged (ops/sense_mix.py)
        if inference_params.cache_position is not None:
            # write position / context lengths on the device (CUDA-graph decode step, utils/generation.py)
            if not self.fused_sense_mix or qk.shape[1] != 1:
                raise RuntimeError("a device-offset decode step needs the fused sense-mix and one position per call")
            kvd = inference_params.key_value_memory_dict
            if "backpack.ctx_k" not in kvd:
                raise RuntimeError("device-offset decoding starts after the prompt pass has allocated the caches")
            b0 = inference_params.batch_size_offset
            b1 = b0 + qk.shape[0]
            k_cache, ids_cache = kvd["backpack.ctx_k"][b0:b1], kvd["backpack.ids"][b0:b1]
            k_cache.index_copy_(1, inference_params.cache_position, qk[:, :, 1])
            ids_cache.index_copy_(1, inference_params.cache_position, input_ids)
            table = self.current_sense_table(force=True)
            out = sense_mix_decode(qk[:, 0, 0], k_cache, ids_cache, table, 0, softmax_scale=scale,
                                   seqlens=inference_params.cache_lengths[b0:b1])
            return out.unsqueeze(1)
        k_cache, ids_cache = self._decode_caches(qk, input_ids, inference_params)
        if offset > 0 and qk.shape[1] != 1:
            raise RuntimeError("after the prompt pass, decoding advances one position per call")
        if self.fused_sense_mix:
            table = self.current_sense_table(force=offset > 0)
            if offset == 0:
                if table is not None:
                    return sense_mix_table(qk, table, input_ids, softmax_scale=scale)
                return sense_mix(qk, self.content_model(input_ids), softmax_scale=scale)
            out = sense_mix_decode(qk[:, 0, 0], k_cache, ids_cache, table, offset + 1, softmax_scale=scale)
            return out.unsqueeze(1)
        if offset == 0:
            content = self.content(input_ids)
            return torch.sum(attn(contextl_hidden_states) @ content, dim=1)
        n = offset + 1
        content = self.content(ids_cache[:, :n])                                         # (b, nv, n, d)
        scores = torch.einsum("bthd,bshd->bhts", qk[:, :, 0], k_cache[:, :n] * scale)   # (b, nv, 1, n)
        alpha = torch.softmax(scores, dim=-1, dtype=qk.dtype)
        return torch.sum(alpha @ content, dim=1)


class BackpackLMHeadModel(BackpackPreTrainedModel, GenerationMixin):

    def __init__(self, config: BackpackConfig, process_group=None, device=None, dtype=None):
        super().__init__(config)
        _no_tp(process_group)
        factory_kwargs = {"device": device, "dtype": dtype}
        self.process_group = None
        self.transformer = BackpackModel(config, **factory_kwargs)
        self.lm_head = nn.Linear(config.n_embd, config.vocab_size, bias=False, **factory_kwargs)
        self.apply(partial(_init_weights, n_layer=config.num_hidden_layers,
                           initializer_range=config.initializer_range))
        self.tie_weights()

    def tie_weights(self):
        self.lm_head.weight = self.transformer.embeddings.word_embeddings.weight

    def graphed_decode_ok(self):
        """Whether a decode step can run with device-side offsets (and so inside a CUDA graph): every operator of the
        step must be one of this library's CUDA kernels or a stream-ordered torch op."""
        cfg = self.config
        return bool(getattr(cfg, "use_flash_attn", False) and self.transformer.fused_sense_mix and not self.training
                    and getattr(cfg, "rotary_emb_fraction", 0.0) == 0.0 and cfg.n_embd // cfg.n_head in (64, 128))

    def forward(self, input_ids, position_ids=None, inference_params=None, num_last_tokens=0):
        """num_last_tokens > 0: project only the last positions to the vocabulary (the generation loop consumes
        logits[:, -1] only, training/src/utils/generation.py:34-44); 0 (the reference's behaviour): all positions."""
        hidden_states = self.transformer(input_ids, position_ids=position_ids, inference_params=inference_params)
        if num_last_tokens > 0:
            hidden_states = hidden_states[:, -num_last_tokens:]
        if getattr(self.config, "fused_bias_fc", False):
            # the tied LM head is a plain GEMM (backpack.py:339-340, 349): same dispatch as every other linear
            return CausalLMOutput(logits=linear(hidden_states, self.lm_head.weight, self.lm_head.bias))
        return CausalLMOutput(logits=self.lm_head(hidden_states))

    @torch.no_grad()
    def token_stats(self, input_ids, targets=None, position_ids=None, num_last_tokens=0):
        """Per-position log-sum-exp, arg-max, max logit (and the logit of `targets`) of the LM head WITHOUT writing
        the logits (ops/lm_head.py): what perplexity evaluation and greedy decoding consume."""
        from ..ops.lm_head import lm_head_stats
        hidden_states = self.transformer(input_ids, position_ids=position_ids)
        if num_last_tokens > 0:
            hidden_states = hidden_states[:, -num_last_tokens:]
            targets = targets[:, -num_last_tokens:] if targets is not None else None
        return lm_head_stats(hidden_states, self.lm_head.weight, targets)

    @torch.no_grad()
    def loss(self, input_ids, labels, ignore_index=-100, reduction="mean"):
        """Next-token cross-entropy (labels already shifted by the caller, as in training/src/tasks/seq.py) through the
        fused LM head: no (batch, seq, vocab) tensor is formed."""
        from ..ops.lm_head import lm_head_cross_entropy
        return lm_head_cross_entropy(self.transformer(input_ids), self.lm_head.weight, labels, ignore_index, reduction)


def flash_config(**kwargs) -> BackpackConfig:
    """The reference's optimised configuration (training/configs/experiment/owt/backpack-small-flash.yaml:10-14,
    training/configs/model/backpack.yaml:11-13) with every fused flag on."""
    base = dict(num_content_vectors=16, vocab_size=50257, activation_function="gelu_new",
                scale_attn_by_inverse_layer_idx=True, reorder_and_upcast_attn=False)
    base.update(kwargs)
    cfg = BackpackConfig(**base)
    cfg.use_flash_attn = True
    cfg.fused_bias_fc = True
    cfg.fused_dense_gelu_dense = True
    cfg.fused_dropout_add_ln = True
    cfg.pad_vocab_size_multiple = 8
    return cfg


def serving_config(**kwargs) -> BackpackConfig:
    """`flash_config` plus the inference-only sense-vector table (`use_sense_table`): in eval mode the content model
    runs once per vocabulary item (at the first forward, or `build_sense_table()`) instead of once per token."""
    cfg = flash_config(**kwargs)
    cfg.use_sense_table = True
    return cfg
