"""CPU (gloo, world_size 2): host-side logic of the row-sharded path -- ownership arithmetic, shard/unshard round trip, and
the partial-softmax exchange protocol (all_reduce MAX of partial maxima, rescale, SUM) restated with torch CPU ops
against the oracle's log-softmax.  The CUDA kernels themselves are covered by tests/test_gpu_sharded.py."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unirec_b200 import sharding


def test_shard_roundtrip_and_ownership():
    full = torch.arange(23 * 4, dtype=torch.float32).view(23, 4)
    for W in (1, 2, 3, 8):
        shards = [sharding.shard_table(full, W, r) for r in range(W)]
        assert [s.shape[0] for s in shards] == [sharding.local_rows_count(23, W, r) for r in range(W)]
        assert torch.equal(sharding.unshard_tables(shards), full)
        ids = torch.arange(23)
        for r in range(W):
            mine = ids[sharding.owner_of(ids, W) == r]
            assert torch.equal(shards[r], full[mine]) and torch.equal(sharding.local_row(mine, W), torch.arange(len(mine)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    V, d, B, N = 64, 8, 4, 9                      # every rank builds the same global problem
    E, U = torch.randn(V, d, generator=g), torch.randn(world * B, d, generator=g)
    ids = torch.randint(0, V, (world * B, N), generator=g)
    s = (E[ids] * U[:, None, :]).sum(-1)
    own = sharding.owner_of(ids, world) == rank
    neg_inf = torch.full_like(s, float('-inf'))
    s_own = torch.where(own, s, neg_inf)
    m = s_own.max(1).values                                        # partial max (may be -inf: no owned entry)
    l = torch.where(own, torch.exp(s - m[:, None]), torch.zeros_like(s)).sum(1)
    l = torch.where(torch.isfinite(m), l, torch.zeros_like(l))
    gm = m.clone()
    dist.all_reduce(gm, op=dist.ReduceOp.MAX)                      # protocol step 1
    scale = torch.where(l > 0, torch.exp(m - gm), torch.zeros_like(l))
    l = l * scale                                                  # protocol step 2 (ur_score_rescale)
    dist.all_reduce(l, op=dist.ReduceOp.SUM)                       # protocol step 3 (reduce_scatter in the product)
    lse = gm + torch.log(l)
    ref = torch.logsumexp(s, dim=1)
    out[rank] = float((lse - ref).abs().max())
    dist.destroy_process_group()


def test_partial_softmax_protocol_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29711, out), nprocs=2, join=True)
    assert len(out) == 2 and max(out.values()) < 1e-5, dict(out)


def _ckpt_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from types import SimpleNamespace
    g = torch.Generator().manual_seed(5)
    V, d = 1003, 8                       # not a multiple of the world size: ranks hold different row counts
    full = torch.randn(V, d, generator=g)
    local = sharding.shard_table(full, world, rank)

    class M(SimpleNamespace):
        def state_dict(self):
            return {'item_embedding.weight': local, 'LayerNorm.weight': torch.ones(d)}
    m = M(shard_world=world, shard_rank=rank, n_items=V)
    sd = sharding.full_state_dict(m)
    ok = True
    if rank == 0:
        ok = torch.equal(sd['item_embedding.weight'], full) and torch.equal(sd['LayerNorm.weight'], torch.ones(d))
    # gather in small chunks too (several send/recv rounds per rank)
    again = sharding.gather_full_table(local, V, world, rank, chunk_rows=100)
    if rank == 0:
        ok = ok and torch.equal(again, full)
    # loading a full (reference-style) checkpoint cuts it back to the local rows; local shards pass through
    back = sharding.localize_state_dict(m, {'item_embedding.weight': full, 'LayerNorm.weight': torch.ones(d)})
    ok = ok and torch.equal(back['item_embedding.weight'], local)
    ok = ok and torch.equal(sharding.localize_state_dict(m, {'item_embedding.weight': local})['item_embedding.weight'], local)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_checkpoint_gather_and_localize_gloo_world2():
    """§8 f4: checkpoints of row-sharded models hold full tables (reference-compatible) and load back into shards."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ckpt_worker, args=(2, 29713, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}, dict(out)


def test_device_loader_gives_every_rank_the_same_number_of_equal_batches():
    """order[rank::world] left ranks with sample counts differing by one -> different step counts / last-batch sizes under
    fixed-shape collectives.  The stream is padded by wrap-around (accelerate's even_batches)."""
    import numpy as np
    from unirec_b200.data.device_loader import DeviceBatchLoader

    class DS:
        return_key_2_index = {'user_id': 0, 'item_id': 1, 'label': 2}

        def __init__(self, n):
            self.user_id = np.arange(1, n + 1, dtype=np.int64)
            self.item_id = np.arange(1, n + 1, dtype=np.int64)

        def __len__(self):
            return len(self.user_id)

    for n, world, bs in [(1001, 8, 16), (64, 8, 8), (17, 4, 5), (5, 8, 2)]:
        for drop_last in (False, True):
            loaders = [DeviceBatchLoader(DS(n), bs, 'cpu', 4, n + 1, n + 1, rank=r, world=world, drop_last=drop_last)
                       for r in range(world)]
            assert len({len(ld) for ld in loaders}) == 1
            assert len({ld._my_count() for ld in loaders}) == 1
            assert loaders[0]._my_count() * world >= n
