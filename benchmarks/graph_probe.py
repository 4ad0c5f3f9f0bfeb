"""Does replaying the forward as one CUDA graph beat the eager launch sequence? (launch gaps between ~130 kernels)"""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, flash_config
from backpacks_flash_attn_b200.utils.weights import name_seeded_
dev = torch.device('cuda', 0)
cfg = flash_config(**bench.SMALL)
model = name_seeded_(BackpackLMHeadModel(cfg).eval()).to(dev, torch.bfloat16)
ids = bench.make_ids(64, 1024).to(dev)
def timeit(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
with torch.inference_mode():
    for _ in range(3): model(ids).logits
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): model(ids).logits
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        out = model(ids).logits
    ref = model(ids).logits
    g.replay(); torch.cuda.synchronize()
    print("graph output equals eager:", torch.equal(out, ref))
    for rep in range(3):
        print("eager %.3f ms | graph %.3f ms" % (timeit(lambda: model(ids).logits), timeit(g.replay)), flush=True)
