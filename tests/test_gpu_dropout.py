"""GPU: counter-based dropout of the fused encoder (csrc/dropout.cuh) -- bit-exact masks vs the numpy Philox restatement,
mask consistency between forward and backward (through the reference-pinned dropout goldens in test_gpu_models.py), keep
rates, eval-mode identity, fresh masks per step under CUDA-graph replay, and the stock SASRec.yaml recipe training."""
import numpy as np
import pytest
import torch

from golden_util import Golden, rel_err
from oracle import philox
from oracle import unirec_oracle as O
from test_host_logic import build_model

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rng(seed, step):
    return torch.tensor([seed, step], dtype=torch.int64, device=DEV)


@pytest.mark.parametrize('p', [0.1, 0.5, 0.9])
def test_mask_kernel_equals_numpy_philox_bit_for_bit(p):
    from unirec_b200 import ops
    seed, step, site, rows, d = 2022, 41, 5, 77, 64
    got = ops.dropout_mask(_rng(seed, step), p, site, rows, d).cpu().numpy()
    assert np.array_equal(got, philox.mask_rows(seed, step, site, p, rows, d))
    # permuted row positions (packed token map) and the compact last-layer map r -> r*L + L-1
    pos = torch.randperm(300, device=DEV)[:rows].to(torch.int32)
    got = ops.dropout_mask(_rng(seed, step), p, site, rows, d, row_pos=pos).cpu().numpy()
    full = philox.mask_rows(seed, step, site, p, 300, d)
    assert np.array_equal(got, full[pos.cpu().numpy()])
    got = ops.dropout_mask(_rng(seed, step), p, site, 9, d, pos_mul=13, pos_add=12).cpu().numpy()
    assert np.array_equal(got, philox.mask_rows(seed, step, site, p, 9 * 13, d)[12::13])
    # flat (attention-probability) indexing
    got = ops.dropout_mask(_rng(seed, step), p, site, 50, 37, flat=True).cpu().numpy().reshape(-1)
    assert np.array_equal(got, philox.mask_flat(seed, step, site, p, 50 * 37))


def test_large_seed_and_64bit_element_index():
    from unirec_b200 import ops
    seed = 0x7A5C3E19F00D1234
    got = ops.dropout_mask(_rng(seed, 3), 0.5, 2, 4, 8, pos_mul=1, pos_add=(1 << 33) // 8).cpu().numpy().reshape(-1)
    assert np.array_equal(got, philox.mask_flat(seed, 3, 2, 0.5, 32, start=(1 << 33)))


def test_dropout_rows_in_place_and_keep_rate():
    from unirec_b200 import ops
    x = torch.ones(4096, 128, device=DEV)
    drop = ops.Drop(_rng(9, 1), 0.3, 0)
    ops.dropout_rows(x, drop)
    keep = float((x > 0).float().mean())
    assert abs(keep - 0.7) < 5e-3
    assert torch.allclose(x[x > 0], torch.tensor(1.0 / 0.7, device=DEV))
    y = torch.ones(4096, 128, device=DEV)
    ops.dropout_rows(y, ops.Drop(_rng(9, 1), 0.0, 0))
    assert torch.equal(y, torch.ones_like(y))       # p = 0: untouched


@pytest.mark.parametrize('name', ['sasrec_softmax_drop', 'sasrec_stock_drop', 'gru_bpr_drop'])
def test_eval_mode_is_identity_and_training_uses_masks(name):
    g = Golden(name)
    model, _ = build_model(g, device=DEV)
    model = model.to(DEV)
    model.load_state_dict(g.params)
    batch = {k: v.to(DEV) for k, v in g.fwd_batch().items()}
    model.eval()
    _, s_eval, u_eval, _ = model(**batch)
    _, so, uo, _ = O.forward(g.model, g.params, g.cfg, **g.fwd_batch())
    assert rel_err(u_eval.cpu(), uo) < 1e-3 and rel_err(s_eval.cpu(), so) < 1e-3
    model.train()
    g.arm(model)
    _, _, u_train, _ = model(**batch, return_loss_only=False)
    assert rel_err(u_train.cpu(), g.user_emb) < 1e-3
    assert rel_err(u_train.cpu(), uo) > 1e-2            # the masks did something
    # a second training forward advances the step: new masks, different output
    _, _, u_next, _ = model(**batch, return_loss_only=False)
    assert rel_err(u_next.cpu(), u_train.cpu()) > 1e-2
    _, _, uo2, _ = O.forward(g.model, g.params, g.cfg, drop=g.drop_masks(1), **g.fwd_batch())
    assert rel_err(u_next.cpu(), uo2) < 1e-3


def test_p0_is_bit_identical_to_dropout_free_build():
    """hidden/attn dropout = 0 takes the mask-free branch of every kernel: train and eval forward agree bit for bit."""
    g = Golden('sasrec_softmax')
    model, _ = build_model(g, device=DEV)
    model = model.to(DEV)
    model.load_state_dict(g.params)
    batch = {k: v.to(DEV) for k, v in g.fwd_batch().items()}
    model.train()
    _, s1, u1, _ = model(**batch, return_loss_only=False)
    model.eval()
    _, s2, u2, _ = model(**batch)
    assert torch.equal(u1, u2) and torch.equal(s1.view(-1), s2.view(-1))


def test_cuda_graph_replay_draws_fresh_masks(tmp_path):
    """The step counter lives on the device and is advanced inside the captured step: replays differ from each other exactly as
    eager steps do, and the graph run matches the eager run step for step."""
    from unirec_b200.facility.accelerator import Accelerator
    from unirec_b200.facility.trainer import Trainer
    from unirec_b200.utils import argument_parser, general
    g = Golden('sasrec_softmax_drop')
    runs = []
    for use_graph in (1, 0):
        args = dict(g.cfg)
        args.update(exp_name='dg', dataset='example', output_path=str(tmp_path), scheduler='none', optimizer='sgd',
                    learning_rate=0.0, cuda_graph=use_graph, epochs=1)
        cfg = argument_parser.parse_arguments(args, argv=[])
        acc = Accelerator()
        cfg['device'] = acc.device
        general.init_seed(3)
        model = general.get_class_instance(g.model, 'unirec_b200/model')(cfg).to(acc.device)
        model.load_state_dict(g.params)
        tr = Trainer(cfg, model, acc)
        g.arm(model)
        batch = {k: v.to(DEV) for k, v in g.fwd_batch().items()}
        runs.append([float(tr.train_step(batch)) for _ in range(6)])       # lr = 0: the loss changes only through the masks
        if use_graph:
            assert all('graph' in v for v in tr._graphs.values())
    graph, eager = runs
    assert max(abs(a - b) for a, b in zip(graph, eager)) < 1e-5 * max(abs(x) for x in eager), (graph, eager)
    assert len({round(x, 5) for x in graph}) == 6
    for it in range(6):
        ref = float(O.forward(g.model, g.params, g.cfg, drop=g.drop_masks(it), **g.fwd_batch())[0])
        assert abs(graph[it] - ref) <= 1e-3 * abs(ref)


def test_stock_sasrec_yaml_trains(tmp_path):
    """config/model/SASRec.yaml unmodified (n_heads 16, hidden/attn dropout 0.5): the loss goes down on a learnable toy task."""
    from unirec_b200.facility.accelerator import Accelerator
    from unirec_b200.facility.trainer import Trainer
    from unirec_b200.utils import argument_parser, general
    cfg = argument_parser.parse_arguments(dict(model='SASRec', dataset='example', exp_name='stock', n_users=50, n_items=400,
                                               train_file_format='user-item', loss_type='softmax', epochs=1, batch_size=256,
                                               learning_rate=1e-2, scheduler='none', output_path=str(tmp_path)), argv=[])
    assert cfg['n_heads'] == 16 and cfg['hidden_dropout_prob'] == 0.5 and cfg['attn_dropout_prob'] == 0.5
    acc = Accelerator()
    cfg['device'] = acc.device
    general.init_seed(11)
    model = general.get_class_instance('SASRec', 'unirec_b200/model')(cfg)
    tr = Trainer(cfg, model, acc)
    L, V, B, K = int(cfg['max_seq_len']), 400, 256, 20
    gen = torch.Generator().manual_seed(5)
    losses = []
    for step in range(200):
        seq = torch.randint(1, V, (B, L), generator=gen)
        target = seq[:, -1] % (V - 1) + 1                               # next item = last item + 1
        neg = torch.randint(1, V, (B, K), generator=gen)
        label = torch.zeros(B, 1 + K, dtype=torch.int32)
        label[:, 0] = 1
        batch = dict(user_id=torch.ones(B, dtype=torch.int64), item_id=torch.cat([target[:, None], neg], 1), label=label,
                     item_seq=seq.to(torch.int32), item_seq_len=torch.full((B,), L, dtype=torch.int64))
        losses.append(float(tr.train_step({k: v.to(DEV) for k, v in batch.items()})))
    assert all(np.isfinite(losses))
    assert np.mean(losses[-10:]) < 0.9 * np.mean(losses[:5]), (losses[:5], losses[-10:])
