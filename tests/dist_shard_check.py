"""Worker for the row-sharded path (run under torch.distributed.run, one process per GPU, or with WORLD_SIZE=1).

Every rank takes its slice of a golden fixture's batch, holds its shard of the golden tables, runs one fused step and
checks: global loss == reference loss; after one step the re-assembled tables and the (replicated) encoder equal one
step of the oracle's dense Adam (identical to lazy Adam on the first step)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main(name='sasrec_softmax', p2p='1'):
    sys.argv = sys.argv[:1]
    from golden_util import Golden, rel_err
    from oracle import unirec_oracle as O
    from unirec_b200 import sharding
    from unirec_b200.facility.optim import FusedOptimizer
    from unirec_b200.utils import argument_parser, general

    W, r = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    if not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29533')
        dist.init_process_group('nccl', rank=r, world_size=W)
    g = Golden(name)
    B = g.batch['item_id'].shape[0]
    B -= B % W                           # fixtures whose batch does not split evenly: the last sample(s) are dropped
    gbatch = {k: v[:B].contiguous() for k, v in g.fwd_batch().items()}
    args = dict(g.cfg)
    args.update(exp_name='shard', dataset='example', table_shard_world=W, table_shard_rank=r, table_shard_force=True,
                shard_p2p=int(p2p))
    cfg = argument_parser.parse_arguments(args, argv=[])
    cfg['device'] = dev
    general.init_seed(2022)
    model = general.get_class_instance(cfg['model'], 'unirec_b200/model')(cfg).to(dev)
    sd = {k: v for k, v in g.params.items()}
    for tname in sharding.SHARDED_TABLES:
        if tname in sd:
            sd[tname] = sharding.shard_table(g.params[tname], W, r)
    model.load_state_dict(sd)
    # ---- evaluation with sharded tables: every rank runs the SAME full batch, owners contribute, ranks sum ----
    from unirec_b200 import ops
    model.eval()
    full = {k: v.to(dev) for k, v in gbatch.items()}
    _, s_eval, u_eval, it_eval = model(**full)
    _, so, uo, _ = O.forward(g.model, g.params, g.cfg, **gbatch)
    assert rel_err(u_eval.cpu(), uo) < 1e-3 and rel_err(s_eval.cpu(), so) < 1e-3, (rel_err(u_eval.cpu(), uo), rel_err(s_eval.cpu(), so))
    assert torch.equal(it_eval.cpu(), g.params['item_embedding.weight'][gbatch['item_id']])          # bit-exact row gather
    target = full['item_id'].view(full['item_id'].shape[0], -1)[:, 0].contiguous()
    counts = model._engine.rank_one_vs_all(u_eval, target, user_id=full.get('user_id'))
    ft = g.params['item_embedding.weight'].to(dev)
    t1 = torch.empty(target.numel(), device=dev)
    c1 = torch.zeros(target.numel(), dtype=torch.int32, device=dev)
    kw = dict(tau=model.tau, item_bias=g.params['item_bias'].to(dev) if model.has_item_bias else None,
              user_bias=g.params['user_bias'].to(dev) if model.has_user_bias else None, user_id=full.get('user_id'))
    ops.rank_target(ft, u_eval, target, t1, **kw)
    ops.rank_count(ft, u_eval, target, t1, c1, **kw)
    ops.rank_exclude(ft, u_eval, target, t1, c1, **kw)
    assert torch.equal(counts, c1), (counts, c1)
    if W > 1:
        full_np = model.forward_all_item_emb()
        assert (torch.from_numpy(full_np) == g.params['item_embedding.weight']).all()
    model.train()
    model._ur_fast_grads = True
    lr = float(cfg['learning_rate'])
    opt = FusedOptimizer(model, 'adam', lr=lr)
    sl = slice(r * (B // W), (r + 1) * (B // W))
    batch = {k: v[sl].contiguous().to(dev) for k, v in gbatch.items()}
    loss = model(**batch)[0]
    opt.zero_grad()
    loss.backward()
    model._engine.sync_dense_grads()
    opt.step()
    torch.cuda.synchronize()
    # oracle: one dense-Adam step on the global batch (identical to lazy Adam on the first step)
    p = O.tie_aliases(g.model, g.cfg, {k: v.clone() for k, v in g.params.items()})
    ref_loss = O.train_step(g.model, p, g.cfg, gbatch, O.DenseAdam(p, lr=lr))
    assert abs(float(loss) - float(ref_loss)) <= 1e-3 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    worst = 0.0
    msd = model.state_dict()
    for k, v in msd.items():
        if k.endswith('key.bias'):
            continue
        if k in sharding.SHARDED_TABLES:
            n_rows = int(getattr(model, sharding.SHARDED_TABLES[k]))
            shards = [torch.empty(sharding.local_rows_count(n_rows, W, q), v.shape[1], device=dev) for q in range(W)]
            for q in range(W):
                if q == r:
                    shards[q].copy_(v)
                dist.broadcast(shards[q], src=q)
            v = sharding.unshard_tables([s.cpu() for s in shards])
        worst = max(worst, rel_err(v.cpu(), p[k]))
    assert worst < 2e-3, worst
    if r == 0:
        print('SHARD_OK world=%d p2p=%d loss=%.6f ref=%.6f max_rel_err=%.2e' % (W, int(model._engine.p2p), float(loss), float(ref_loss), worst))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main(*sys.argv[1:])
