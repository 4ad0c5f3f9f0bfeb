import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, flash_config
from backpacks_flash_attn_b200.utils.weights import name_seeded_
dev = torch.device('cuda', 0)
cfg = flash_config(**bench.SMALL)
model = name_seeded_(BackpackLMHeadModel(cfg).eval()).to(dev, torch.bfloat16)
ids_host = bench.make_ids(64, 1024).contiguous().pin_memory()
ids_dev = ids_host.to(dev)
last_host = torch.empty((2, 64, cfg.vocab_size), dtype=torch.bfloat16).pin_memory()
copy_stream = torch.cuda.Stream(device=dev)
def resident(n):
    for _ in range(n): model(ids_dev).logits
def e2e(n, mode):
    pending = None
    main = torch.cuda.current_stream()
    for i in range(n):
        x = ids_host.to(dev, non_blocking=True)
        logits = model(x).logits
        if mode == 'nocopy':
            del logits; continue
        last_dev = logits[:, -1].contiguous(); del logits
        ready = torch.cuda.Event(); ready.record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            last_host[i % 2].copy_(last_dev, non_blocking=True)
            last_dev.record_stream(copy_stream)
            done = torch.cuda.Event(); done.record(copy_stream)
        if mode == 'pipelined':
            if pending is not None: pending.synchronize()
            pending = done
        elif mode == 'sync':
            done.synchronize()
    torch.cuda.synchronize()
def timeit(fn, *a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(*a); torch.cuda.synchronize(); return (time.perf_counter() - t0) / a[0] * 1e3
with torch.inference_mode():
    resident(3)
    for rep in range(2):
        print('resident %.2f | e2e pipelined %.2f | nocopy %.2f | sync %.2f | resident %.2f' % (
            timeit(resident, 10), timeit(e2e, 10, 'pipelined'), timeit(e2e, 10, 'nocopy'), timeit(e2e, 10, 'sync'), timeit(resident, 10)), flush=True)
