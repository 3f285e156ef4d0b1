"""CPU: host-side logic of the drop-in boundary -- config precedence, registry, seed-exact initialisation and
state_dict names (vs the reference goldens), trainer control helpers, loud failure without CUDA."""
import os
import sys

import pytest
import torch

from golden_util import CASES, Golden
from unirec_b200.utils import argument_parser, general


def _cfg(g, **over):
    args = dict(g.cfg)
    args.update(exp_name='t', dataset='example')
    args.update(over)
    cfg = argument_parser.parse_arguments(args, argv=[])
    cfg['device'] = torch.device('cpu')
    return cfg


def build_model(g, device='cpu', **over):
    cfg = _cfg(g, **over)
    cfg['device'] = torch.device(device)
    general.init_seed(2022)
    return general.get_class_instance(cfg['model'], 'unirec_b200/model')(cfg), cfg


@pytest.mark.parametrize('name', CASES)
def test_seeded_init_is_bit_identical_to_reference(name):
    g = Golden(name)
    model, _ = build_model(g)
    sd = model.state_dict()
    assert set(sd) == set(g.params)
    for k in sd:
        assert torch.equal(sd[k], g.params[k]), k


def test_config_precedence_and_cmd_args(tmp_path):
    f = tmp_path / 'extra.yaml'
    f.write_text('embedding_size: 48\nlearning_rate: 0.5\n')
    cfg = argument_parser.parse_arguments({'model': 'SASRec', 'dataset': 'example', 'learning_rate': 0.25},
                                          argv=['--config_file', str(f), '--n_layers=3', '--bogus_flag=1', '--scheduler=none'])
    assert cfg['n_heads'] == 16                 # model yaml over base
    assert cfg['embedding_size'] == 48          # --config_file over yaml tree
    assert cfg['n_layers'] == 3                 # command line over files
    assert cfg['learning_rate'] == 0.25         # args dict over everything
    assert cfg['scheduler'] == 'reduce'         # 'none' dropped
    assert 'bogus_flag' not in cfg
    assert cfg['cmd_args']['n_layers'] == 3


def test_registry_resolves_reference_spelling():
    for name in ('SASRec', 'GRU', 'AvgHist', 'SVDPlusPlus', 'MF'):
        assert general.get_class_instance(name, 'unirec_b200/model').__name__ == name
        assert general.get_class_instance(name, 'unirec/model').__name__ == name
    with pytest.raises(ValueError):
        general.get_class_instance('NoSuchModel', 'unirec_b200/model')


def test_unsupported_options_fail_loudly():
    g = Golden('mf_bpr')
    with pytest.raises(ValueError):
        build_model(g, loss_type='bce')
    with pytest.raises(ValueError):
        build_model(g, distance_type='cosine')


def test_product_path_has_no_cpu_fallback():
    g = Golden('mf_bpr')
    model, _ = build_model(g)
    model.train()
    with pytest.raises(RuntimeError, match='CUDA'):
        model(**g.fwd_batch())


def test_early_stopping_rule():
    from unirec_b200.facility.trainer import Trainer
    best, step, stop, upd = Trainer.early_stopping(0.5, None, 1, max_step=2)
    assert (best, step, stop, upd) == (0.5, 0, False, True)
    best, step, stop, upd = Trainer.early_stopping(0.4, best, step, max_step=2)
    assert (best, step, stop, upd) == (0.5, 1, False, False)
    best, step, stop, upd = Trainer.early_stopping(0.4, best, 2, max_step=2)
    assert stop and not upd
    assert Trainer.early_stopping(0.1, 0.9, 7, max_step=0) == (0.9, 7, False, True)


def test_rank_metrics():
    from unirec_b200.facility.evaluation import RankEvaluator
    ev = RankEvaluator("['hit@1;3', 'ndcg@3', 'mrr', 'group_auc']")
    res = ev.metrics_from_ranks(torch.tensor([0., 2., 5.], dtype=torch.float64), 11)
    assert res['hit@1'] == pytest.approx(1 / 3) and res['hit@3'] == pytest.approx(2 / 3)
    assert res['ndcg@3'] == pytest.approx((1.0 + 0.5) / 3)
    assert res['mrr'] == pytest.approx((1 + 1 / 3 + 1 / 6) / 3)
    assert res['group_auc'] == pytest.approx((1.0 + 0.8 + 0.5) / 3)


def test_rank_metrics_match_reference_evaluator_golden():
    """§8 f3: hit / ndcg / mrr / group_auc from ranks equal the reference's OnePositiveEvaluator on the same score matrices
    (fixture written by oracle/make_eval_golden.py from unirec/facility/evaluation/onepos.py:104-175)."""
    import os
    import numpy as np
    from unirec_b200.facility.evaluation import RankEvaluator
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_eval', 'onepos_metrics.npz'))
    metrics_str = bytes(z['metrics_str']).decode()
    ev = RankEvaluator(metrics_str)
    for tag in ('a', 'b'):
        scores = torch.from_numpy(z['scores_' + tag])
        rank = (scores[:, 1:] > scores[:, :1]).sum(1).double()          # one_vs_k rank, as RankEvaluator._ranks computes it
        res = ev.metrics_from_ranks(rank, scores.shape[1])
        ref = {k.split('/', 1)[1]: float(z[k]) for k in z.files if k.startswith('metric_%s/' % tag)}
        assert set(ref) == set(res), (sorted(ref), sorted(res))
        for k, v in ref.items():
            assert res[k] == pytest.approx(v, rel=1e-6, abs=1e-9), (tag, k, res[k], v)


def test_hot_path_switches_have_documented_defaults():
    """The B200-specific config keys (DESIGN.md sections 4-5) parse from base.yaml with the defaults the bench runs with, and
    command-line / dict overrides reach the config like any reference key."""
    cfg = argument_parser.parse_arguments({'model': 'SASRec', 'dataset': 'example'}, argv=[])
    assert cfg['gemm_precision'] == 'tf32x3' and int(cfg['pack_sequences']) == 1 and int(cfg['trim_last_layer']) == 1
    assert cfg.get('table_update', 'sparse') == 'sparse'
    cfg = argument_parser.parse_arguments({'model': 'SASRec', 'dataset': 'example', 'pack_sequences': 0},
                                          argv=['--gemm_precision=fp32', '--trim_last_layer=0'])
    assert cfg['gemm_precision'] == 'fp32' and int(cfg['pack_sequences']) == 0 and int(cfg['trim_last_layer']) == 0


def test_install_as_unirec_alias_resolves_reference_imports():
    """code written against the reference package (`from unirec.main import main`,
    `get_class_instance(name, 'unirec/model')`) resolves to this package after `install_as_unirec()`.  Runs in a subprocess so the
    alias does not leak into the other tests."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); import unirec_b200; unirec_b200.install_as_unirec();"
            "from unirec.utils import general, argument_parser; from unirec.main import main;"
            "import unirec.data.dataset, unirec.facility.trainer;"
            "names = ['SASRec', 'GRU', 'AvgHist', 'SVDPlusPlus', 'MF'];"
            "mods = [general.get_class_instance(n, 'unirec/model').__module__ for n in names];"
            "assert all(m.startswith('unirec_b200.model.') for m in mods), mods;"
            "assert callable(main.run); print('ALIAS_OK')" % root)
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120, cwd='/tmp')
    assert res.returncode == 0 and 'ALIAS_OK' in res.stdout, res.stdout + res.stderr


def test_reference_written_checkpoint_loads_into_drop_in_classes():
    """§8 f4: a checkpoint written by the reference (its model class, its dict layout: trainer.py:389-398; fixture from
    oracle/make_ckpt_golden.py) is rebuilt by load_model_freely into the unirec_b200 class with identical weights."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_eval', 'reference_checkpoint_mf.pth')
    cpt = torch.load(path, map_location='cpu', weights_only=False)
    assert {'config', 'cur_epoch', 'cur_step', 'best_valid_score', 'state_dict', 'optimizer', 'scheduler'} <= set(cpt)
    model, cfg = general.load_model_freely(path, device=torch.device('cpu'))
    assert type(model).__module__.startswith('unirec_b200.') and type(model).__name__ == 'MF'
    sd = model.state_dict()
    assert set(sd) == set(cpt['state_dict'])
    for k, v in cpt['state_dict'].items():
        assert torch.equal(sd[k], v), k
