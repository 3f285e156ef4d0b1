"""GPU: the drop-in model classes + trainer pieces against the reference goldens (tests/golden/*.npz, produced by the
reference's own classes) and the CPU oracle.  Bars: loss / scores / user_emb within 1e-3 relative fp32 (north_star)."""
import pytest
import torch

from golden_util import CASES, Golden, rel_err
from oracle import unirec_oracle as O
from test_host_logic import build_model

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-3


def cuda_model(g, **over):
    model, cfg = build_model(g, device=DEV, **over)
    model = model.to(DEV)
    model.load_state_dict(g.params)
    return model, cfg


def _traj_tol(name):
    """3-step Adam trajectories: 2e-3.  The pad-edge fixtures hold an empty-history sample whose attention logits all sit at
    s - 10000, i.e. on the 1e-3 grid of fp32 -- the reference's own output for it is quantised, summation-order differences flip
    grid points and Adam turns them into fractions of lr (the CPU oracle itself is 6e-4 off the reference there)."""
    return 6e-3 if (name.endswith('padedges') or name == 'sasrec_softmax_drop') else 2e-3


def to_dev(batch):
    return {k: v.to(DEV) for k, v in batch.items()}


@pytest.mark.parametrize('name', CASES)
def test_forward_matches_reference(name):
    g = Golden(name)
    model, _ = cuda_model(g)
    model.train()
    g.arm(model)            # dropout fixtures: draw the mask set the reference ran with (no-op otherwise)
    loss, scores, user_emb, items_emb = model(**to_dev(g.fwd_batch()), return_loss_only=False)
    assert abs(float(loss) - float(g.loss)) <= TOL * abs(float(g.loss))
    assert rel_err(scores.cpu(), g.scores) < TOL
    assert rel_err(user_emb.cpu(), g.user_emb) < TOL
    assert torch.equal(items_emb.cpu(), O.gather_rows(g.params['item_embedding.weight'], g.batch['item_id']))   # bit-exact
    g.arm(model)
    loss_vec = model(**to_dev(g.fwd_batch()), reduction=False)[0]
    assert rel_err(loss_vec.detach().cpu(), g.loss_vec) < TOL
    model.eval()
    none, s2, u2, _ = model(**to_dev(g.fwd_batch()))
    if g.drop_step is None:
        assert none is None and rel_err(s2.cpu(), g.scores) < TOL and rel_err(u2.cpu(), g.user_emb) < TOL
    else:       # eval mode = dropout off: the oracle without masks
        _, so, uo, _ = O.forward(g.model, g.params, g.cfg, **g.fwd_batch())
        assert none is None and rel_err(s2.cpu(), so) < TOL and rel_err(u2.cpu(), uo) < TOL


@pytest.mark.parametrize('name', CASES)
def test_autograd_grads_match_reference_dense_mode(name):
    """`table_update: dense` + loss.backward(): every parameter's .grad equals the reference's dense gradient."""
    g = Golden(name)
    model, _ = cuda_model(g, table_update='dense')
    model.train()
    g.arm(model)
    loss = model(**to_dev(g.fwd_batch()))[0]
    loss.backward()
    scale = max(float(v.abs().max()) for v in g.grads.values())
    for k, p in model.named_parameters():
        ref = g.grads[k]
        got = p.grad.cpu() if p.grad is not None else torch.zeros_like(ref)
        err = float((got.double() - ref.double()).abs().max())
        assert err <= TOL * max(float(ref.abs().max()), 1e-2 * scale), (k, err)


@pytest.mark.parametrize('name', CASES)
def test_dense_mode_trajectory_matches_reference_adam(name):
    """3 steps of the fused optimizer in exact-dense mode == the reference loop with torch.optim.Adam."""
    from unirec_b200.facility.optim import FusedOptimizer
    g = Golden(name)
    model, cfg = cuda_model(g, table_update='dense')
    model.train()
    model._ur_fast_grads = True
    opt = FusedOptimizer(model, 'adam', lr=float(cfg['learning_rate']), weight_decay=float(cfg['weight_decay']))
    batch = to_dev(g.fwd_batch())
    g.arm(model)            # the step counter then advances by one per training forward, like the fixture's mask sets
    for ref_loss in g.traj_loss:
        loss = model(**batch)[0]
        opt.zero_grad()
        loss.backward()
        opt.step()
        assert abs(float(loss) - ref_loss) <= TOL * abs(ref_loss)
    sd = model.state_dict()
    for k, ref in g.traj_params.items():
        if k.endswith('key.bias'):
            continue     # analytically-zero gradient: Adam amplifies rounding noise (see test_oracle_golden)
        assert rel_err(sd[k].cpu(), ref) < _traj_tol(name), k


@pytest.mark.parametrize('name', CASES)
def test_sparse_mode_trajectory_matches_lazy_oracle(name):
    """Default mode: row-sparse table update == oracle LazyRowAdam (dense Adam restricted to touched rows)."""
    from unirec_b200.facility.optim import FusedOptimizer
    g = Golden(name)
    model, cfg = cuda_model(g)
    model.train()
    model._ur_fast_grads = True
    lr = float(cfg['learning_rate'])
    opt = FusedOptimizer(model, 'adam', lr=lr)
    p = O.tie_aliases(g.model, g.cfg, {k: v.clone() for k, v in g.params.items()})
    oopt = O.LazyRowAdam(p, lr=lr)
    b = g.batch
    seq_table = 'item_dst_embedding.weight' if (g.model == 'SVDPlusPlus' or (g.model == 'AvgHist' and g.cfg.get('asymmetric', True))) \
        else 'item_embedding.weight'
    touched = {'item_embedding.weight': [b['item_id'].reshape(-1)]}
    if g.model != 'MF':
        touched.setdefault(seq_table, []).append(b['item_seq'].reshape(-1).long())
    if 'user_embedding.weight' in g.params:
        touched['user_embedding.weight'] = [b['user_id']]
    touched = {k: torch.cat(v) for k, v in touched.items()}
    batch = to_dev(g.fwd_batch())
    g.arm(model)
    for it in range(3):
        loss = model(**batch)[0]
        opt.zero_grad()
        loss.backward()
        opt.step()
        ref_loss, _, _, grads = O.loss_and_grads(g.model, p, g.cfg, g.fwd_batch(), drop=g.drop_masks(it))
        oopt.step(p, grads, touched)
        assert abs(float(loss) - float(ref_loss)) <= TOL * abs(float(ref_loss))
    sd = model.state_dict()
    for k, ref in p.items():
        if k.endswith('key.bias'):
            continue
        assert rel_err(sd[k].cpu(), ref) < _traj_tol(name), k


def test_grad_clipping_matches_oracle():
    from unirec_b200.facility.optim import FusedOptimizer
    g = Golden('sasrec_softmax')
    model, cfg = cuda_model(g)
    model.train()
    model._ur_fast_grads = True
    opt = FusedOptimizer(model, 'sgd', lr=0.5)
    opt.max_grad_norm = 0.05
    loss = model(**to_dev(g.fwd_batch()))[0]
    opt.zero_grad()
    loss.backward()
    opt.step()
    _, _, _, grads = O.loss_and_grads(g.model, g.params, g.cfg, g.fwd_batch())
    total = O.clip_grad_norm(grads, 0.05)
    assert float(total) > 0.05      # the clip is active in this fixture
    sd = model.state_dict()
    for k, ref in g.params.items():
        if k.endswith('key.bias'):
            continue     # analytically-zero gradient (rounding noise only)
        assert rel_err(sd[k].cpu(), ref - 0.5 * grads[k]) < 2e-3, k


def test_nan_loss_skips_update_without_sync():
    from unirec_b200.facility.optim import FusedOptimizer
    g = Golden('mf_bpr')
    model, cfg = cuda_model(g)
    model.train()
    model._ur_fast_grads = True
    opt = FusedOptimizer(model, 'adam', lr=0.1)
    with torch.no_grad():
        model.user_embedding.weight[int(g.batch['user_id'][0])] = float('nan')
    before = {k: v.clone() for k, v in model.state_dict().items()}
    loss = model(**to_dev(g.fwd_batch()))[0]
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert torch.isnan(loss)
    for k, v in model.state_dict().items():
        assert torch.equal(torch.nan_to_num(v), torch.nan_to_num(before[k])), k
    assert int((model._engine.rowgrad(model.item_embedding.weight).head != -1).sum()) == 0


def test_trainer_fit_loop_and_checkpoint(tmp_path):
    """Trainer.fit through the reference-shaped loop on a synthetic in-memory dataset (MF+BPR = BASELINE config 1 shape)."""
    from unirec_b200.facility.accelerator import Accelerator
    from unirec_b200.facility.trainer import Trainer
    from unirec_b200.utils import argument_parser, general

    class DS(torch.utils.data.Dataset):
        return_key_2_index = {'user_id': 0, 'item_id': 1, 'label': 2}

        def __init__(self, n, U, V, K, seed):
            gg = torch.Generator().manual_seed(seed)
            self.u = torch.randint(1, U, (n,), generator=gg)
            self.i = torch.randint(1, V, (n, 1 + K), generator=gg)
            self.i[:, 0] = (self.u * 7) % (V - 1) + 1          # learnable signal
            self.l = torch.zeros(n, 1 + K, dtype=torch.int32)
            self.l[:, 0] = 1

        def __len__(self):
            return len(self.u)

        def __getitem__(self, k):
            return self.u[k], self.i[k], self.l[k]

    cfg = argument_parser.parse_arguments(dict(model='MF', dataset='example', exp_name='fit', n_users=200, n_items=300,
                                               embedding_size=64, loss_type='bpr', train_file_format='user-item', epochs=3,
                                               batch_size=256, learning_rate=0.05, scheduler='none', early_stop=0,
                                               output_path=str(tmp_path), metrics="['hit@5','group_auc']",
                                               key_metric='group_auc'), argv=[])
    acc = Accelerator()
    cfg['device'] = acc.device
    general.init_seed(7)
    model = general.get_class_instance('MF', 'unirec_b200/model')(cfg)
    tr = Trainer(cfg, model, acc)
    tr.reset_evaluator('user-item', 'one_vs_k')
    train = torch.utils.data.DataLoader(DS(4096, 200, 300, 1, 1), batch_size=256)
    valid = torch.utils.data.DataLoader(DS(512, 200, 300, 9, 2), batch_size=256)
    tr.fit(train, valid, save_model=True)
    res = tr.evaluate(valid, load_best_model=True)
    assert res['group_auc'] > 0.8, res
    model2, _ = general.load_model_freely(tr.saved_model_file, device=acc.device)
    assert set(model2.state_dict()) == set(model.state_dict())


@pytest.mark.parametrize('name', ['sasrec_softmax', 'sasrec_softmax_padedges', 'sasrec_bpr_nopos_padedges', 'sasrec_softmax_d128',
                                  'sasrec_softmax_drop', 'sasrec_stock_drop', 'sasrec_h16_d64'])
@pytest.mark.parametrize('pack,trim', [(0, 1), (1, 0), (0, 0)])
def test_sasrec_packing_and_trimming_do_not_change_results(name, pack, trim):
    """pack_sequences / trim_last_layer only remove dead work: every combination must reproduce the reference goldens (the default
    1/1 combination is what every other test runs)."""
    g = Golden(name)
    model, _ = cuda_model(g, table_update='dense', pack_sequences=pack, trim_last_layer=trim)
    model.train()
    g.arm(model)            # dropout masks are indexed by original positions: packing / trimming must not move them
    loss, scores, user_emb, _ = model(**to_dev(g.fwd_batch()), return_loss_only=False)
    assert abs(float(loss) - float(g.loss)) <= TOL * abs(float(g.loss))
    assert rel_err(scores.cpu(), g.scores) < TOL and rel_err(user_emb.cpu(), g.user_emb) < TOL
    loss.backward()
    scale = max(float(v.abs().max()) for v in g.grads.values())
    for k, p in model.named_parameters():
        ref = g.grads[k]
        err = float((p.grad.cpu().double() - ref.double()).abs().max())
        assert err <= TOL * max(float(ref.abs().max()), 1e-2 * scale), (k, err)


@pytest.mark.parametrize('trim', [1, 0])
def test_sasrec_d128_3xtf32_tensor_core_path_meets_fp32_bar(trim):
    """gemm_precision=tf32x3 (the bench default): tcgen05 GEMMs with the hi/lo operand split -> the fp32 parity bar (1e-3) holds,
    with and without the last-layer trimming (dead rows of the last encoder layer not computed)."""
    g = Golden('sasrec_softmax_d128')
    model, _ = cuda_model(g, table_update='dense', gemm_precision='tf32x3', trim_last_layer=trim)
    model.train()
    loss, scores, user_emb, _ = model(**to_dev(g.fwd_batch()), return_loss_only=False)
    assert abs(float(loss) - float(g.loss)) <= 1e-4 * abs(float(g.loss))
    assert rel_err(scores.cpu(), g.scores) < 1e-4 and rel_err(user_emb.cpu(), g.user_emb) < 1e-4
    loss.backward()
    scale = max(float(v.abs().max()) for v in g.grads.values())
    for k, p in model.named_parameters():
        ref = g.grads[k]
        err = float((p.grad.cpu().double() - ref.double()).abs().max())
        assert err <= 1e-3 * max(float(ref.abs().max()), 1e-2 * scale), (k, err)


def test_sasrec_d128_tf32_tensor_core_path():
    """gemm_precision=tf32 routes the encoder GEMMs through tcgen05; parity bar for reduced-precision modes is 1e-2."""
    g = Golden('sasrec_softmax_d128')
    model, _ = cuda_model(g, table_update='dense', gemm_precision='tf32')
    model.train()
    loss, scores, user_emb, _ = model(**to_dev(g.fwd_batch()), return_loss_only=False)
    assert abs(float(loss) - float(g.loss)) <= 1e-2 * abs(float(g.loss))
    assert rel_err(scores.cpu(), g.scores) < 1e-2 and rel_err(user_emb.cpu(), g.user_emb) < 1e-2
    loss.backward()
    scale = max(float(v.abs().max()) for v in g.grads.values())
    for k, p in model.named_parameters():
        ref = g.grads[k]
        err = float((p.grad.cpu().double() - ref.double()).abs().max())
        assert err <= 1e-2 * max(float(ref.abs().max()), 1e-2 * scale), (k, err)


@pytest.mark.parametrize('mode', [1, 2])
@pytest.mark.parametrize('name', ['sasrec_softmax', 'gru_bpr', 'sasrec_softmax_d128'])
def test_trainer_overlapped_table_update_matches_serial(name, mode, tmp_path):
    """Trainer.train_step with the early-linked, two-phase table update (target-only rows updated on a side stream during
    the encoder backward) must land on the same parameters and optimizer state as the serial update."""
    from unirec_b200.facility.accelerator import Accelerator
    from unirec_b200.facility.trainer import Trainer
    from unirec_b200.utils import argument_parser, general
    g = Golden(name)
    results = []
    for overlap in (mode, 0):
        args = dict(g.cfg)
        args.update(exp_name='ovl', dataset='example', output_path=str(tmp_path), scheduler='none', optimizer='adam',
                    learning_rate=1e-2, overlap_table_update=overlap, epochs=1)
        cfg = argument_parser.parse_arguments(args, argv=[])
        acc = Accelerator()
        cfg['device'] = acc.device
        general.init_seed(3)
        model = general.get_class_instance(g.model, 'unirec_b200/model')(cfg).to(acc.device)
        model.load_state_dict(g.params)
        tr = Trainer(cfg, model, acc)
        # Adam's eps raised from 1e-8: with the default, gradient elements at the round-off floor (|g| ~ 1e-9, sign decided by the
        # order of float atomics in the split-K GEMMs) become +-lr/10 parameter moves and make the comparison flaky (seen 1 in 5 runs)
        tr.optimizer.param_groups[0]['eps'] = 1e-5
        assert (model._engine.overlap_hook is not None) == bool(overlap)
        losses = []
        for step in range(4):
            batch = to_dev(g.fwd_batch())
            if step % 2:        # vary the batch: shift the ids (kept in range) so the row sets change between steps
                V = int(g.cfg['n_items'])
                batch['item_id'] = (batch['item_id'] % (V - 1)) + 1
            losses.append(float(tr.train_step(batch)))
        torch.cuda.synchronize()
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        opt_sd = tr.optimizer.state_dict()
        results.append((losses, sd, opt_sd))
    (l1, s1, o1), (l0, s0, o0) = results
    assert max(abs(a - b) for a, b in zip(l1, l0)) < 1e-5 * max(abs(x) for x in l0)
    for k in s0:
        if k.endswith('key.bias'):       # mathematically zero gradient (softmax shift invariance): Adam amplifies its round-off
            continue
        assert rel_err(s1[k], s0[k]) < 1e-4, (k, rel_err(s1[k], s0[k]))
    for tname in o0['tables']:
        for mv in o0['tables'][tname]:
            assert rel_err(o1['tables'][tname][mv].cpu(), o0['tables'][tname][mv].cpu()) < 1e-4, (tname, mv)
    assert int(o1['step']) == int(o0['step']) == 4


@pytest.mark.parametrize('name', ['sasrec_softmax', 'gru_bpr', 'mf_bpr', 'sasrec_softmax_d128'])
def test_trainer_cuda_graph_step_matches_eager(name, tmp_path):
    """Trainer.train_step replays the captured CUDA graph from the third step of a batch signature on: losses, parameters and
    optimizer state must match the eager steps (same kernels, same order; atomics make runs agree to round-off, not bitwise)."""
    from unirec_b200.facility.accelerator import Accelerator
    from unirec_b200.facility.trainer import Trainer
    from unirec_b200.utils import argument_parser, general
    g = Golden(name)
    results = []
    for use_graph in (1, 0):
        args = dict(g.cfg)
        args.update(exp_name='cg', dataset='example', output_path=str(tmp_path), scheduler='none', optimizer='adam',
                    learning_rate=1e-2, cuda_graph=use_graph, epochs=1)
        cfg = argument_parser.parse_arguments(args, argv=[])
        acc = Accelerator()
        cfg['device'] = acc.device
        general.init_seed(3)
        model = general.get_class_instance(g.model, 'unirec_b200/model')(cfg).to(acc.device)
        model.load_state_dict(g.params)
        tr = Trainer(cfg, model, acc)
        tr.optimizer.param_groups[0]['eps'] = 1e-5      # (see test_trainer_overlapped_table_update_matches_serial)
        losses = []
        for step in range(6):
            batch = to_dev(g.fwd_batch())
            if step % 2:
                V = int(g.cfg['n_items'])
                batch['item_id'] = (batch['item_id'] % (V - 1)) + 1
            losses.append(float(tr.train_step(batch)))
        torch.cuda.synchronize()
        assert bool(getattr(tr, '_graphs', {})) == bool(use_graph)
        if use_graph:
            assert all('graph' in v for v in tr._graphs.values())
        results.append((losses, {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}))
    (l1, s1), (l0, s0) = results
    # same kernels in the same order; split-K / bias-gradient reductions use float atomics, so runs agree to round-off only
    assert max(abs(a - b) for a, b in zip(l1, l0)) < 1e-5 * max(abs(x) for x in l0), (l1, l0)
    for k in s0:
        if k.endswith('key.bias'):
            continue
        assert rel_err(s1[k], s0[k]) < 1e-4, (k, rel_err(s1[k], s0[k]))


def test_frozen_parameters_do_not_move():
    """Trainer.load_model(freeze) sets requires_grad = False (reference trainer.py:383-386): the fused optimizer must leave those
    parameters alone -- dense ones (runs of the flat buffer) and whole tables."""
    from unirec_b200.facility.optim import FusedOptimizer
    g = Golden('sasrec_softmax')
    model, cfg = cuda_model(g)
    model.train()
    model._ur_fast_grads = True
    frozen = ['item_embedding.weight', 'trm_encoder.layer.0.feed_forward.dense_1.weight', 'LayerNorm.bias']
    for n, p in model.named_parameters():
        if n in frozen:
            p.requires_grad = False
    opt = FusedOptimizer(model, 'adam', lr=0.05, weight_decay=0.01)
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for _ in range(2):
        loss = model(**to_dev(g.fwd_batch()))[0]
        opt.zero_grad()
        loss.backward()
        opt.step()
    after = model.state_dict()
    for k in before:
        if k in frozen:
            assert torch.equal(after[k], before[k]), k
    assert not torch.equal(after['trm_encoder.layer.0.feed_forward.dense_2.weight'], before['trm_encoder.layer.0.feed_forward.dense_2.weight'])
    assert not torch.equal(after['position_embedding.weight'], before['position_embedding.weight'])
    head = model._engine.rowgrad(model.item_embedding.weight).head
    assert head is None or int((head != -1).sum()) == 0


@pytest.mark.parametrize('name', ['gru_bpr', 'gru_softmax_h32', 'gru_d256_h256_L100', 'gru_bpr_drop'])
def test_gru_persistent_recurrence_matches_reference(name):
    """gru_persistent=1: the whole recurrence in one launch per direction (csrc/gru.cu) reproduces the reference goldens (forward,
    dense gradients), incl. the BASELINE c3 recurrence shape d = h = 256, L = 100."""
    g = Golden(name)
    model, _ = cuda_model(g, table_update='dense', gru_persistent=1)
    model.train()
    g.arm(model)
    loss, scores, user_emb, _ = model(**to_dev(g.fwd_batch()), return_loss_only=False)
    assert abs(float(loss) - float(g.loss)) <= TOL * abs(float(g.loss))
    assert rel_err(scores.cpu(), g.scores) < TOL and rel_err(user_emb.cpu(), g.user_emb) < TOL
    loss.backward()
    scale = max(float(v.abs().max()) for v in g.grads.values())
    for k, p in model.named_parameters():
        ref = g.grads[k]
        err = float((p.grad.cpu().double() - ref.double()).abs().max())
        assert err <= TOL * max(float(ref.abs().max()), 1e-2 * scale), (k, err)


def test_reference_checkpoint_scores_on_the_cuda_path():
    """The reference-written checkpoint (tests/golden_eval/reference_checkpoint_mf.pth) evaluated by the CUDA path reproduces the scores
    the reference model computed before saving."""
    import os
    from unirec_b200.utils import general
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_eval', 'reference_checkpoint_mf.pth')
    model, _ = general.load_model_freely(path, device=torch.device(DEV))
    probe = torch.load(path, map_location='cpu', weights_only=False)['_probe']
    model.eval()
    _, scores, _, _ = model(user_id=probe['user_id'].to(DEV), item_id=probe['item_id'].to(DEV))
    assert rel_err(scores.cpu(), probe['scores']) < TOL
