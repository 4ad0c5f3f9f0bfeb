"""world_size-2 gloo tests of the batch-sharding plumbing (the N>1 path of bench.py)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from backpacks_flash_attn_b200 import parallel


def test_shard_range_partitions_the_batch():
    for gb in (1, 7, 64, 512, 513):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def _worker(rank, world, port, diverge, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    try:
        r, lr, w = parallel.init_distributed("gloo")
        assert (r, w) == (rank, world)
        from backpacks_flash_attn_b200.models.backpack import BackpackConfig, BackpackLMHeadModel
        from backpacks_flash_attn_b200.utils.weights import name_seeded_
        cfg = BackpackConfig(num_content_vectors=4, n_embd=32, n_head=2, n_layer=1, n_positions=16, vocab_size=96,
                             pad_vocab_size_multiple=8)
        model = name_seeded_(BackpackLMHeadModel(cfg).eval())
        if diverge and rank == 1:
            with torch.no_grad():
                model.lm_head.weight[0, 0] += 1.0
        try:
            parallel.assert_replicas_match(model)
            matched = True
        except RuntimeError:
            matched = False
        ids = torch.randint(0, 96, (6, 16), generator=torch.Generator().manual_seed(1234))
        mine = parallel.shard_batch(ids, rank, world)
        with torch.no_grad():
            logits = model(mine).logits
            full = model(ids).logits
        nxt = parallel.gather_next_tokens(logits[:, -1])
        ok = torch.equal(nxt, full[:, -1].argmax(-1)) if not diverge else True
        t = parallel.max_over_ranks(float(rank + 1), "cpu")
        tok = parallel.sum_over_ranks(float(mine.numel()), "cpu")
        parallel.barrier()
        q.put((rank, matched, ok, t, tok))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("diverge", [False, True])
def test_two_rank_gloo_sharded_forward(diverge):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + int(diverge)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, diverge, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, matched, ok, t, tok in res:
        assert matched == (not diverge)
        assert ok
        assert t == 2.0                 # max over ranks
        assert tok == 6 * 16            # every token processed exactly once


def _train_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    try:
        parallel.init_distributed("gloo")
        import torch.nn.functional as F
        from backpacks_flash_attn_b200.models.backpack import BackpackConfig, BackpackLMHeadModel
        from backpacks_flash_attn_b200.utils.weights import name_seeded_
        cfg = BackpackConfig(num_content_vectors=4, n_embd=32, n_head=2, n_layer=1, n_positions=16, vocab_size=96,
                             pad_vocab_size_multiple=8, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
        ids = torch.randint(0, 96, (6, 16), generator=torch.Generator().manual_seed(4321))

        def grads_of(batch, scale):
            model = name_seeded_(BackpackLMHeadModel(cfg).train())
            logits = model(batch).logits
            loss = F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), batch[:, 1:].reshape(-1), reduction="sum")
            (loss * scale).backward()
            return model

        n_tok = ids.shape[0] * (ids.shape[1] - 1)
        mine = parallel.shard_batch(ids, rank, world)
        model = grads_of(mine, world / n_tok)             # mean over ranks of (world / n) * local sum = global mean
        calls = parallel.allreduce_gradients(model, bucket_bytes=4096)
        whole = grads_of(ids, 1.0 / n_tok)                # the same step on the un-sharded batch
        worst = max((p.grad - w.grad).abs().max().item() for p, w in zip(model.parameters(), whole.parameters()))
        q.put((rank, calls, worst))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_matches_the_unsharded_step():
    """Data-parallel training step: sharded batch + bucketed gradient all-reduce == the step on the whole batch."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, 29631, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, calls, worst in res:
        assert calls >= 2                 # several buckets at this bucket size
        assert worst < 1e-5, worst
