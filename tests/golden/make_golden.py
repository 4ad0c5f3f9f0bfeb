"""Generate the golden fixtures in this directory FROM THE REAL REFERENCE.

Run once in the build container (``/root/reference`` is not present on the GPU box):

    python tests/golden/make_golden.py

It imports the unmodified reference (pure-PyTorch path, fp32, CPU) with the import shims of
SURVEY.md §8c -- none of which touches a reference file:
  1/2. ``transformers.generation.{GreedySearch,Sample}DecoderOnlyOutput`` aliases (renamed upstream);
  3.   ``backpack.FusedDense = nn.Linear`` (``FusedDense`` subclasses ``nn.Linear`` and its forward is
       ``F.linear``; it is ``None`` here because ``fused_dense_lib`` cannot be built);
  4.   an empty stand-in module named ``rotary_emb`` so ``flash_attn/layers/rotary.py`` imports
       (only its pure-torch ``apply_rotary_emb_torch`` is used).
Outputs: ``micro_model.npz`` (Backpack-Micro end to end, recipe weights), ``small_model.npz``
(Backpack-Small, a few slices) and ``ops.npz`` (operator-level cases).
"""
import os
import sys
import types
import zlib

import numpy as np
import torch

REF = os.environ.get("BP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "training"))
    import transformers.generation as tg
    tg.GreedySearchDecoderOnlyOutput = tg.GenerateDecoderOnlyOutput
    tg.SampleDecoderOnlyOutput = tg.GenerateDecoderOnlyOutput
    sys.modules.setdefault("rotary_emb", types.ModuleType("rotary_emb"))
    import src.models.backpack as bp
    bp.FusedDense = torch.nn.Linear
    return bp


def seed_weights(model):
    """The RNG-order independent recipe of SURVEY.md §8c, applied to the reference model."""
    with torch.no_grad():
        for name, p in sorted(dict(model.named_parameters()).items()):
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
            w = torch.randn(p.shape, generator=g)
            if p.dim() == 2:
                p.copy_(w * p.shape[1] ** -0.5)
            elif name.endswith(("ln_0.weight", "norm1.weight", "norm2.weight")):
                p.copy_(1 + 0.1 * w)
            else:
                p.copy_(0.02 * w)


def build(bp, n_embd, n_head, n_layer, n_positions):
    cfg = bp.BackpackConfig(num_content_vectors=16, n_embd=n_embd, n_head=n_head, n_layer=n_layer,
                            n_positions=n_positions, vocab_size=50257, reorder_and_upcast_attn=False,
                            scale_attn_by_inverse_layer_idx=True, pad_vocab_size_multiple=8)
    model = bp.BackpackLMHeadModel(cfg).eval()
    seed_weights(model)
    return model


def run_model(model, ids):
    t = model.transformer
    with torch.no_grad():
        ctx_h = t.gpt2_model(ids)
        alpha = t.contextualization_attn(ctx_h)
        content = t.content_model(ids)
        hid = torch.sum(alpha @ content, dim=1)
        assert torch.equal(hid, t(ids))
        logits = model(ids).logits
    return ctx_h, alpha, content, hid, logits


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    bp = import_reference()

    # ---- Backpack-Micro, ids (2,128): SURVEY.md §8c ----
    model = build(bp, 384, 6, 6, 512)
    ids = torch.randint(0, 50257, (2, 128), generator=torch.Generator().manual_seed(1234))
    ctx_h, alpha, content, hid, logits = run_model(model, ids)
    np.savez(os.path.join(HERE, "micro_model.npz"),
             ids=ids.numpy(), ctx_h=ctx_h.numpy(), hid=hid.numpy(),
             alpha_0_3=alpha[0, 3].numpy(), alpha_1_15=alpha[1, 15].numpy(),
             alpha_rowsum=alpha.sum(-1).numpy(),
             alpha_upper_max=np.float32(alpha.triu(1).abs().max().item()),
             content_0_0=content[0, 0].numpy(), content_1_15=content[1, 15].numpy(),
             content_strides=np.array(content.stride()),
             logits_head=logits[:, :, :64].numpy(), logits_last=logits[1, 127].numpy(),
             argmax=logits.argmax(-1).numpy(),
             mean_abs=np.array([ctx_h.abs().mean(), content.abs().mean(), hid.abs().mean(),
                                logits.abs().mean()], dtype=np.float64))
    print("micro:", hid[0, 0, :4].tolist(), logits[0, 0, :4].tolist())

    # ---- Backpack-Small, ids (1,256): SURVEY.md Appendix B ----
    model = build(bp, 768, 12, 12, 1024)
    ids2 = torch.randint(0, 50257, (1, 256), generator=torch.Generator().manual_seed(1234))
    ctx_h, alpha, content, hid, logits = run_model(model, ids2)
    np.savez(os.path.join(HERE, "small_model.npz"),
             ids=ids2.numpy(), ctx_h_rows=ctx_h[0, ::32].numpy(), hid=hid.numpy(),
             alpha_0_5_200=alpha[0, 5, 200].numpy(), alpha_0_15_255=alpha[0, 15, 255].numpy(),
             content_0_3_100=content[0, 3, 100].numpy(),
             logits_last_head=logits[0, 255, :256].numpy(), argmax=logits.argmax(-1).numpy(),
             mean_abs=np.array([ctx_h.abs().mean(), content.abs().mean(), hid.abs().mean(),
                                logits.abs().mean()], dtype=np.float64))
    print("small:", hid[0, 0, :4].tolist(), logits[0, 255, :4].tolist())
    del model

    # ---- operator-level cases ----
    from flash_attn.modules.mha import SelfAttention
    from flash_attn.modules.mlp import Mlp
    from flash_attn.layers.rotary import apply_rotary_emb_torch
    import torch.nn.functional as F
    from functools import partial
    out = {}
    g = torch.Generator().manual_seed(7)
    # eager attention, fp32 and bf16, two scales (mha.py:195-224)
    qkv = torch.randn(2, 96, 3, 4, 32, generator=g)
    out["attn_qkv"] = qkv.numpy()
    with torch.no_grad():
        out["attn_causal"] = SelfAttention(causal=True)(qkv).numpy()
        out["attn_causal_scale"] = SelfAttention(causal=True, softmax_scale=0.0625)(qkv).numpy()
        out["attn_full"] = SelfAttention(causal=False)(qkv).numpy()
        out["attn_causal_bf16"] = SelfAttention(causal=True)(qkv.bfloat16()).float().numpy()
    # ContextSelfAttn + sense sum (backpack.py:94-122, 313)
    csa = bp.ContextSelfAttn(8, 128)
    with torch.no_grad():
        csa.Wqkv.weight.copy_(torch.randn(256, 128, generator=g) * 128 ** -0.5)
        csa.Wqkv.bias.copy_(torch.randn(256, generator=g) * 0.02)
        h = torch.randn(2, 64, 128, generator=g)
        a = csa(h)
        content = torch.randn(2, 64, 8, 128, generator=g).transpose(1, 2)  # transposed view, like :276
        out["ctx_w"], out["ctx_b"], out["ctx_h"] = csa.Wqkv.weight.numpy(), csa.Wqkv.bias.numpy(), h.numpy()
        out["ctx_alpha"] = a.numpy()
        out["ctx_content_bsnd"] = content.transpose(1, 2).contiguous().numpy()
        out["ctx_sense_sum"] = torch.sum(a @ content, dim=1).numpy()
        out["ctx_alpha_bf16"] = csa.bfloat16()(h.bfloat16()).float().numpy()
    # Mlp with tanh-GELU (mlp.py:13-30, gpt.py:87-89)
    m = Mlp(64, hidden_features=256, activation=partial(F.gelu, approximate="tanh"))
    with torch.no_grad():
        x = torch.randn(5, 7, 64, generator=g)
        out["mlp_w1"], out["mlp_b1"] = m.fc1.weight.numpy().copy(), m.fc1.bias.numpy().copy()
        out["mlp_w2"], out["mlp_b2"] = m.fc2.weight.numpy().copy(), m.fc2.bias.numpy().copy()
        out["mlp_x"], out["mlp_y"] = x.numpy(), m(x).numpy()
    # rotary (rotary.py:18-28)
    x = torch.randn(2, 33, 3, 64, generator=g)
    ang = torch.rand(33, 16, generator=g) * 6.28
    out["rot_x"], out["rot_cos"], out["rot_sin"] = x.numpy(), ang.cos().numpy(), ang.sin().numpy()
    out["rot_y"] = apply_rotary_emb_torch(x, ang.cos(), ang.sin()).numpy()
    np.savez(os.path.join(HERE, "ops.npz"), **out)
    print("ops:", sorted(out))


if __name__ == "__main__":
    main()
