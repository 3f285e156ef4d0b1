"""CPU: the claim behind `pack_sequences` / `trim_last_layer`, checked on the oracle's restatement of the
reference encoder: positions that csrc/pack.cu drops are DEAD -- whatever the encoder holds there cannot change the user embedding or
any parameter gradient.  The test overwrites the hidden states of the dead positions with large random values after the embedding
LayerNorm and after every encoder layer, and compares against the untouched computation on the reference goldens."""
import pytest
import torch

from golden_util import Golden
from oracle import unirec_oracle as O

CASES = ['sasrec_softmax', 'sasrec_softmax_padedges', 'sasrec_bpr_nopos_padedges', 'sasrec_bpr_nopos_bias']


def live_mask(item_seq):
    """csrc/pack.cu: real items + position L-1; every position of a sequence without any real item."""
    real = item_seq > 0
    keep = real.clone()
    keep[:, -1] = True
    keep |= ~real.any(1, keepdim=True)
    return keep


def user_emb_with_junk(p, cfg, item_seq, junk_scale, gen):
    """O.sasrec_user_emb with the hidden states of the dead positions replaced by noise between all stages."""
    eps = float(cfg['layer_norm_eps'])
    causal = bool(cfg.get('use_position_emb', True))
    dead = ~live_mask(item_seq)

    def poison(x):
        if junk_scale == 0:
            return x
        noise = torch.randn(x.shape, generator=gen, dtype=x.dtype) * junk_scale
        return torch.where(dead[:, :, None], noise, x)

    x = O.gather_rows(p['item_embedding.weight'], item_seq)
    if causal:
        x = x + p['position_embedding.weight'][:item_seq.shape[1]][None]
    x = poison(O.layer_norm(x, p['LayerNorm.weight'], p['LayerNorm.bias'], eps))
    mask = O.sasrec_attention_mask(item_seq, causal, x.dtype)
    n_layers = int(cfg['n_layers'])
    for i in range(n_layers):
        pre = 'trm_encoder.layer.%d.' % i
        x = O.multi_head_attention(x, mask, p, pre + 'multi_head_attention.', int(cfg['n_heads']), eps)
        x = O.feed_forward(x, p, pre + 'feed_forward.', cfg['hidden_act'], eps)
        if i < n_layers - 1:
            x = poison(x)
    return x[:, -1, :]


@pytest.mark.parametrize('name', CASES)
def test_dead_positions_cannot_reach_the_output_or_the_gradients(name):
    g = Golden(name)
    item_seq = g.batch['item_seq']
    assert (~live_mask(item_seq)).any(), 'fixture has no dead position'
    outs = []
    for scale in (0.0, 50.0):
        p = {k: v.detach().clone().double().requires_grad_(True) for k, v in g.params.items()}
        u = user_emb_with_junk(p, g.cfg, item_seq, scale, torch.Generator().manual_seed(1))
        u.square().sum().backward()          # any function of the user embedding: the loss only sees the tower through it
        outs.append((u.detach(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)).clone() for k, v in p.items()}))
    (u0, g0), (u1, g1) = outs
    assert torch.equal(u0, u1)
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k
    # and the untouched restatement is the reference (float64 here vs the fp32 golden; the empty-history sample of the pad-edge
    # fixtures has all logits at s - 10000, which fp32 quantises to a 1e-3 grid -- hence the wider bar there)
    tol = 5e-3 if name.endswith('padedges') else 1e-5
    assert float((u0.float() - g.user_emb).abs().max()) <= tol * float(g.user_emb.abs().max())


def test_live_mask_rules():
    seq = torch.tensor([[0, 0, 3, 4], [0, 5, 0, 0], [0, 0, 0, 0], [7, 0, 8, 9]], dtype=torch.int32)
    keep = live_mask(seq)
    assert keep.tolist() == [[False, False, True, True], [False, True, False, True], [True, True, True, True],
                             [True, False, True, True]]
