"""Small forward + backward cases of every kernel of the library, meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python benchmarks/sanitizer_cases.py
    compute-sanitizer --tool synccheck python benchmarks/sanitizer_cases.py

Shapes are tiny (the tools slow a kernel down by one to two orders of magnitude) but cover ragged tails, the varlen
path, dropout and both storage types.  Results are only checked for finiteness here -- parity lives in tests/.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from backpacks_flash_attn_b200.flash_attn_interface import (flash_attn_unpadded_qkvpacked_func,  # noqa: E402
                                                            flash_attn_unpadded_func)
from backpacks_flash_attn_b200.ops.sense_mix import sense_mix  # noqa: E402
from backpacks_flash_attn_b200.ops.layer_norm import dropout_add_layer_norm  # noqa: E402
from backpacks_flash_attn_b200.ops.fused_dense import FusedDenseGeluDense, FusedDense  # noqa: E402
from backpacks_flash_attn_b200.losses.cross_entropy import CrossEntropyLoss  # noqa: E402

dev = "cuda"


def finite(name, *ts):
    torch.cuda.synchronize()
    ok = all(bool(torch.isfinite(t.float()).all()) for t in ts if t is not None)
    print(("PASS " if ok else "FAIL ") + name, flush=True)
    if not ok:
        raise SystemExit(1)


def attention(d, seqlens, causal, dtype, p=0.0):
    torch.manual_seed(0)
    h = 2
    total = sum(seqlens)
    cu = torch.tensor([0] + list(torch.tensor(seqlens).cumsum(0)), dtype=torch.int32, device=dev)
    qkv = torch.randn(total, 3, h, d, device=dev).to(dtype).requires_grad_()
    out = flash_attn_unpadded_qkvpacked_func(qkv, cu, max(seqlens), p, causal=causal)
    out.backward(torch.randn_like(out))
    finite(f"attention d={d} seqlens={seqlens} causal={causal} {dtype} p={p}", out, qkv.grad)


def cross_attention(d):
    torch.manual_seed(1)
    h, sq, sk, b = 2, 100, 230, 2
    q = torch.randn(b * sq, h, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(b * sk, h, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(b * sk, h, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    cq = torch.arange(0, (b + 1) * sq, sq, dtype=torch.int32, device=dev)
    ck = torch.arange(0, (b + 1) * sk, sk, dtype=torch.int32, device=dev)
    out = flash_attn_unpadded_func(q, k, v, cq, ck, sq, sk, 0.0, causal=False)
    out.backward(torch.randn_like(out))
    finite(f"cross attention d={d}", out, q.grad, k.grad, v.grad)


def sense(b, s, nv, d):
    torch.manual_seed(2)
    qk = torch.randn(b, s, 2, nv, d // nv, device=dev).to(torch.bfloat16).requires_grad_()
    content = (torch.randn(b, s, nv, d, device=dev).to(torch.bfloat16) * 0.5).requires_grad_()
    out = sense_mix(qk, content.transpose(1, 2))
    out.backward(torch.randn_like(out))
    finite(f"sense_mix b={b} s={s} nv={nv} d={d}", out, qk.grad, content.grad)


def layer_norm(rows, cols, p=0.0):
    torch.manual_seed(3)
    x0 = torch.randn(rows, cols, device=dev, dtype=torch.bfloat16, requires_grad=True)
    res = torch.randn(rows, cols, device=dev, dtype=torch.float32, requires_grad=True)
    w = torch.randn(cols, device=dev, dtype=torch.bfloat16, requires_grad=True)
    b = torch.randn(cols, device=dev, dtype=torch.bfloat16, requires_grad=True)
    z, r = dropout_add_layer_norm(x0, res, w, b, p, 1e-5, prenorm=True, residual_in_fp32=True)
    (z.float().sum() + r.sum()).backward()
    finite(f"layer norm {rows}x{cols} p={p}", z, r, x0.grad, res.grad, w.grad, b.grad)


def mlp(rows, n):
    torch.manual_seed(4)
    m = FusedDenseGeluDense(n, 4 * n, n, device=dev, dtype=torch.bfloat16)
    lin = FusedDense(n, 3 * n, device=dev, dtype=torch.bfloat16)
    x = torch.randn(rows, n, device=dev, dtype=torch.bfloat16, requires_grad=True)
    y = lin(m(x))
    y.backward(torch.randn_like(y))
    finite(f"fused dense {rows}x{n}", y, x.grad, m.fc1.weight.grad, m.fc1.bias.grad, lin.weight.grad, lin.bias.grad)


def xent(rows, vocab):
    torch.manual_seed(5)
    logits = torch.randn(rows, vocab, device=dev, dtype=torch.bfloat16, requires_grad=True)
    labels = torch.randint(0, vocab, (rows,), device=dev)
    labels[::7] = -100
    loss = CrossEntropyLoss()(logits, labels)
    loss.backward()
    finite(f"cross entropy {rows}x{vocab}", loss, logits.grad)


if __name__ == "__main__":
    attention(64, [256, 256], True, torch.bfloat16)
    attention(64, [200, 77, 130], True, torch.bfloat16)
    attention(64, [200, 77], False, torch.float16)
    attention(128, [130, 256], True, torch.bfloat16)
    attention(40, [129], True, torch.bfloat16)
    attention(64, [256, 100], True, torch.bfloat16, p=0.1)
    cross_attention(64)
    sense(1, 256, 16, 768)
    sense(2, 200, 4, 768)
    layer_norm(300, 768)
    layer_norm(65, 1024)
    layer_norm(77, 768, p=0.1)
    mlp(300, 256)
    mlp(200, 128)       # fewer than 256 rows: the backward recomputes the pre-activation
    xent(100, 50264)
    xent(33, 1000)
    print("sanitizer cases: all finite")
