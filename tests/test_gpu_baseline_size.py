"""GPU: parity at BASELINE.json sizes (the toy-sized goldens never exercise the multi-wave score kernel, the packed
encoder with ~26k live tokens, split-K weight-gradient GEMMs or million-row row lists).  The CUDA path runs the exact configuration;
the checker is the CPU oracle port (pinned to the reference by tests/golden) on the same seeded inputs and the same parameters.
Bars (north_star): loss / scores / user_emb within 1e-3 relative; parameters after one row-sparse step vs the oracle's lazy Adam."""
import pytest
import torch

from oracle import unirec_oracle as O
from unirec_b200.utils import argument_parser, general

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _build(model, **kw):
    args = dict(model=model, dataset='example', exp_name='size', train_file_format='user-item', scheduler='none', optimizer='adam',
                learning_rate=1e-3, hidden_dropout_prob=0.0, attn_dropout_prob=0.0, dropout_prob=0.0)
    args.update(kw)
    cfg = argument_parser.parse_arguments(args, argv=[])
    cfg['device'] = torch.device(DEV)
    general.init_seed(2022)
    m = general.get_class_instance(model, 'unirec_b200/model')(cfg).to(DEV)
    return m, cfg


def _batch(V, U, B, K, L, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    seq = torch.randint(1, V, (B, L), generator=g)
    seq = torch.where(torch.arange(L)[None, :] >= (L - lens)[:, None], seq, torch.zeros_like(seq)).to(torch.int32)
    label = torch.zeros(B, 1 + K, dtype=torch.int32)
    label[:, 0] = 1
    return dict(user_id=torch.randint(1, U, (B,), generator=g), item_id=torch.randint(1, V, (B, 1 + K), generator=g), label=label,
                item_seq=seq, item_seq_len=lens)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _check_step(model_name, cfg_kw, V, U, B, K, L, seq_table='item_embedding.weight'):
    from unirec_b200.facility.optim import FusedOptimizer
    model, cfg = _build(model_name, n_items=V, n_users=U, max_seq_len=L, **cfg_kw)
    params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    batch = _batch(V, U, B, K, L, 7)
    dbatch = {k: v.to(DEV) for k, v in batch.items()}
    model.train()
    loss, scores, user_emb, _ = model(**dbatch, return_loss_only=False)
    ocfg = {k: v for k, v in cfg.items() if isinstance(v, (int, float, str, bool))}
    p = O.tie_aliases(model_name, ocfg, {k: v.clone() for k, v in params.items()})
    ref_loss, ref_scores, ref_user, grads = O.loss_and_grads(model_name, p, ocfg, batch)
    assert abs(float(loss) - float(ref_loss)) <= 1e-3 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    assert rel(user_emb.cpu(), ref_user) < 1e-3
    assert rel(scores.cpu(), ref_scores) < 1e-3
    # one row-sparse optimizer step through the million-entry row lists.  SGD with lr = 1 makes the parameter change the gradient
    # itself (Adam would turn round-off on near-zero gradient elements into +-lr flips): touched rows and every dense parameter move
    # by -grad within 1e-3 of the gradient scale, untouched rows stay bit-identical.
    model._ur_fast_grads = True
    opt = FusedOptimizer(model, 'sgd', lr=1.0)
    loss2 = model(**dbatch)[0]
    opt.zero_grad()
    loss2.backward()
    opt.step()
    torch.cuda.synchronize()
    touched = {'item_embedding.weight': batch['item_id'].reshape(-1)}
    touched[seq_table] = torch.cat([touched.get(seq_table, torch.zeros(0, dtype=torch.int64)), batch['item_seq'].reshape(-1).long()])
    sd = model.state_dict()
    gscale = max(float(g.abs().max()) for g in grads.values())
    for k, g in grads.items():
        got = sd[k].detach().cpu()
        delta = (got - params[k]).double()
        err = float((delta + g.double()).abs().max())
        assert err <= 1e-3 * max(float(g.abs().max()), 1e-2 * gscale), (k, err, float(g.abs().max()))
        if k in touched:
            untouched = torch.ones(got.shape[0], dtype=torch.bool)
            untouched[torch.unique(touched[k])] = False
            assert torch.equal(got[untouched], params[k][untouched]), k


@pytest.mark.parametrize('prec', ['tf32x3', 'fp32'])
def test_c2_sasrec_exact_configuration_vs_oracle(prec):
    """BASELINE configs[1]: SASRec d=128, 2 layers, 2 heads, L=50, V=1M, softmax K=256, B=1024."""
    _check_step('SASRec', dict(embedding_size=128, hidden_size=128, n_layers=2, n_heads=2, inner_size=512, loss_type='softmax',
                               hidden_act='swish', layer_norm_eps=1e-10, use_position_emb=1, gemm_precision=prec),
                V=1_000_000, U=100_000, B=1024, K=256, L=50)


@pytest.mark.parametrize('persistent', [0, 1])
def test_c3_gru_configuration_vs_oracle(persistent):
    """BASELINE configs[2]: GRU d = h = 256, L=100, BPR K=5, B=2048.  V is 1M here instead of 5M: the CPU checker materialises dense
    [V, d] gradients and Adam state (5M x 256 fp32 x 4 copies = 20 GB of host memory); the row arithmetic does not depend on V."""
    _check_step('GRU', dict(embedding_size=256, hidden_size=256, loss_type='bpr', gemm_precision='tf32x3', gru_persistent=persistent),
                V=1_000_000, U=100_000, B=2048, K=5, L=100)


def test_headline_score_kernel_at_full_size_vs_float64():
    """The metric's own workload: B=1024, N=1+1024, d=128 over a 10M-row table (5.1 GB): multi-wave launch of the cp.async.bulk ring
    kernel.  Every sample's loss term / dLoss/du and the dLoss/ds of sampled rows against float64 on the gathered rows."""
    from unirec_b200 import ops
    torch.manual_seed(11)
    V, d, B, N = 10_000_000, 128, 1024, 1025
    table = torch.randn(V, d, device=DEV) * 0.05
    table[0] = 0
    user = torch.randn(B, d, device=DEV)
    ids = torch.randint(1, V, (B, N), device=DEV)
    ids[5, 7] = 0
    ids[9, 3] = ids[9, 2]
    label = torch.zeros(B, N, dtype=torch.int32, device=DEV)
    label[:, 0] = 1
    n_pos = torch.zeros(1, device=DEV)
    ops.count_positive(label, n_pos)
    scores = torch.empty(B, N, device=DEV)
    loss_vec = torch.empty(B, device=DEV)
    dscore = torch.empty(B, N, device=DEV)
    grad_user = torch.empty(B, d, device=DEV)
    ops.score_loss(table, user, ids, 'softmax', label=label, norm_dev=n_pos, scores=scores, loss_vec=loss_vec, dscore=dscore,
                   grad_user=grad_user)
    rows = ops.gather_rows(table, ids)                                        # bit-exact gather (tests/test_gpu_ops.py)
    s = torch.einsum('bnd,bd->bn', rows.double(), user.double())
    lse = torch.logsumexp(s, 1)
    assert rel(scores, s) < 1e-5
    assert rel(loss_vec, lse - s[:, 0]) < 1e-5
    p = torch.softmax(s, 1)
    y = torch.zeros_like(p)
    y[:, 0] = 1
    ds = (p - y) / B
    assert float((dscore.double() - ds).abs().max()) < 1e-5 * float(ds.abs().max())
    gu = torch.einsum('bn,bnd->bd', ds, rows.double())
    assert rel(grad_user, gu) < 1e-4
