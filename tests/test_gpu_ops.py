"""GPU: every CUDA entry point, called through the C ABI, against the CPU oracle on the same seeded inputs.
Bars (BASELINE.json north_star): bit-exact for index gather; fp32 kernels within 1e-3 relative (measured ~1e-6)."""
import math

import pytest
import torch

from oracle import unirec_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize('d', [16, 32, 64, 128, 256])
@pytest.mark.parametrize('idt', [torch.int32, torch.int64])
def test_gather_rows_bit_exact(d, idt):
    from unirec_b200 import ops
    g = gen(d)
    table = torch.randn(1000, d, generator=g)
    idx = torch.randint(0, 1000, (7, 13), generator=g).to(idt)
    out = ops.gather_rows(table.to(DEV), idx.to(DEV))
    assert torch.equal(out.cpu(), O.gather_rows(table, idx))
    empty = ops.gather_rows(table.to(DEV), torch.zeros(0, dtype=idt, device=DEV))
    assert empty.shape == (0, d)


def test_gather_rows_large_table_property():
    """Full-size property test (10M x 128 table, 1M random rows): gathered row r must equal a table whose row i is
    filled with a function of i -- no oracle needed at this size."""
    from unirec_b200 import ops
    V, d, n = 10_000_000, 128, 1 << 20
    table = (torch.arange(V, device=DEV, dtype=torch.float32) % 65536).unsqueeze(1).expand(V, d).contiguous()
    idx = torch.randint(0, V, (n,), device=DEV)
    out = ops.gather_rows(table, idx)
    assert torch.equal(out[:, 0], (idx % 65536).float()) and torch.equal(out[:, -1], out[:, 0])


def test_scatter_add_rows_matches_dense_embedding_backward():
    from unirec_b200 import ops
    g = gen(1)
    V, d, B, N = 50, 32, 6, 9
    idx = torch.randint(0, V, (B, N), generator=g)
    src = torch.randn(B, d, generator=g)
    coef = torch.randn(B, N, generator=g)
    grad = torch.zeros(V, d, device=DEV)
    ops.scatter_add_rows(grad, idx.to(DEV), src.to(DEV), src_group=N, coef=coef.to(DEV), coef_group=1, pad_id=0)
    ref = torch.zeros(V, d)
    ref.index_add_(0, idx.reshape(-1), (coef.unsqueeze(-1) * src.unsqueeze(1)).reshape(-1, d))
    ref[0] = 0
    assert rel(grad, ref) < 1e-5


@pytest.mark.parametrize('d,L', [(32, 5), (128, 50), (64, 501)])
def test_pool_sum(d, L):
    from unirec_b200 import ops
    g = gen(2)
    V, U, B = 300, 40, 9
    E, Ut = torch.randn(V, d, generator=g), torch.randn(U, d, generator=g)
    E[0] = 0
    seq = torch.randint(0, V, (B, L), generator=g).to(torch.int32)
    ln = torch.randint(1, L + 1, (B,), generator=g)
    uid = torch.randint(1, U, (B,), generator=g)
    cfg = {'user_sequence_alpha': 0.5}
    out = ops.pool_sum_fwd(E.to(DEV), seq.to(DEV), ln.to(DEV), 0.5)
    assert rel(out, O.avghist_user_emb({'item_dst_embedding.weight': E}, cfg, seq, ln)) < 1e-5
    out2 = ops.pool_sum_fwd(E.to(DEV), seq.to(DEV), ln.to(DEV), 0.5, Ut.to(DEV), uid.to(DEV))
    ref2 = O.svdpp_user_emb({'item_dst_embedding.weight': E, 'user_embedding.weight': Ut}, cfg, uid, seq, ln)
    assert rel(out2, ref2) < 1e-5


@pytest.mark.parametrize('d', [32, 128, 256])
@pytest.mark.parametrize('with_pos', [True, False])
def test_seq_prep_ln_fwd_bwd(d, with_pos):
    from unirec_b200 import ops
    g = gen(3)
    V, B, L, eps = 200, 5, 7, 1e-10
    E = torch.randn(V, d, generator=g)
    P = torch.randn(L + 1, d, generator=g) if with_pos else None
    gam, bet = torch.randn(d, generator=g), torch.randn(d, generator=g)
    seq = torch.randint(0, V, (B, L), generator=g).to(torch.int32)
    dY = torch.randn(B, L, d, generator=g)
    Er, gr, br = E.clone().requires_grad_(), gam.clone().requires_grad_(), bet.clone().requires_grad_()
    Pr = P.clone().requires_grad_() if with_pos else None
    x = O.gather_rows(Er, seq)
    if with_pos:
        x = x + Pr[:L][None]
    x.retain_grad()
    y = O.layer_norm(x, gr, br, eps)
    y.backward(dY)
    Y = torch.empty(B * L, d, device=DEV)
    mean, rstd = torch.empty(B * L, device=DEV), torch.empty(B * L, device=DEV)
    Pd = P.to(DEV) if with_pos else None
    ops.seq_prep_ln_fwd(E.to(DEV), Pd, gam.to(DEV), bet.to(DEV), eps, seq.to(DEV), Y, mean, rstd)
    assert rel(Y.view(B, L, d), y) < 1e-5
    dX = torch.empty(B * L, d, device=DEV)
    dg, db = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    dP = torch.zeros(L + 1, d, device=DEV) if with_pos else None
    ops.seq_prep_ln_bwd(E.to(DEV), Pd, gam.to(DEV), seq.to(DEV), mean, rstd, dY.to(DEV).view(B * L, d), dX, dg, db, dP)
    assert rel(dX.view(B, L, d), x.grad) < 1e-4
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4
    if with_pos:
        assert rel(dP, Pr.grad) < 1e-4


@pytest.mark.parametrize('d', [32, 128, 512])
def test_add_ln_fwd_bwd(d):
    from unirec_b200 import ops
    g = gen(4)
    rows, eps = 37, 1e-10
    X, R = torch.randn(rows, d, generator=g), torch.randn(rows, d, generator=g)
    gam, bet = torch.randn(d, generator=g), torch.randn(d, generator=g)
    dY = torch.randn(rows, d, generator=g)
    Z = (X + R).requires_grad_()
    gr, br = gam.clone().requires_grad_(), bet.clone().requires_grad_()
    y = O.layer_norm(Z, gr, br, eps)
    y.backward(dY)
    Xd = X.to(DEV)
    Y = torch.empty(rows, d, device=DEV)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    ops.add_ln_fwd(Xd, R.to(DEV), gam.to(DEV), bet.to(DEV), eps, Y, mean, rstd)
    assert rel(Y, y) < 1e-5 and rel(Xd, Z) < 1e-6
    dZ = torch.empty(rows, d, device=DEV)
    dg, db = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    ops.add_ln_bwd(Xd, gam.to(DEV), mean, rstd, dY.to(DEV), dZ, dg, db)
    assert rel(dZ, Z.grad) < 1e-4 and rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4


@pytest.mark.parametrize('M,N,K', [(300, 128, 128), (51, 384, 64), (1000, 512, 128), (129, 132, 260)])
@pytest.mark.parametrize('act', [None, 'swish', 'gelu', 'relu'])
def test_gemm_nt_bias_act(M, N, K, act):
    from unirec_b200 import ops
    g = gen(5)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.1, torch.randn(N, generator=g)
    C, pre = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), C, M, N, K, transB=True, bias=b.to(DEV), act=act, preact=pre)
    ref_pre = O.linear(A.double(), W.double(), b.double())
    ref = O.activation(ref_pre, act) if act else ref_pre
    assert rel(pre, ref_pre) < 1e-5 and rel(C, ref) < 1e-5


def test_gemm_nn_tn_accumulate_and_colsum():
    from unirec_b200 import ops
    g = gen(6)
    M, N, K = 5000, 96, 132
    dY, W, X = torch.randn(M, N, generator=g), torch.randn(N, K, generator=g), torch.randn(M, K, generator=g)
    dX = torch.ones(M, K, device=DEV)
    ops.gemm(dY.to(DEV), W.to(DEV), dX, M, K, N, accumulate=True)
    assert rel(dX, 1.0 + dY.double() @ W.double()) < 1e-5
    dW = torch.zeros(N, K, device=DEV)
    ops.gemm(dY.to(DEV), X.to(DEV), dW, N, K, M, transA=True, lda=N, accumulate=True)     # split-K path
    assert rel(dW, dY.double().t() @ X.double()) < 1e-5
    db = torch.zeros(N, device=DEV)
    ops.colsum_accum(dY.to(DEV), M, N, db)
    assert rel(db, dY.double().sum(0)) < 1e-5


@pytest.mark.parametrize('L,H,dh', [(8, 2, 16), (50, 2, 64), (11, 4, 16), (200, 2, 32), (33, 1, 128)])
@pytest.mark.parametrize('causal', [True, False])
def test_attention_fwd_bwd(L, H, dh, causal):
    from unirec_b200 import ops
    g = gen(7)
    B, d = 3, H * dh
    qkv = torch.randn(B, L, 3 * d, generator=g)
    seq = torch.randint(1, 100, (B, L), generator=g).to(torch.int32)
    seq[0, : L // 2] = 0
    seq[1, : L - 1] = 0
    live = (seq > 0)
    # gradient only arrives at real positions (padded query rows never reach the loss); the kernels skip them exactly
    dctx = torch.randn(B, L, d, generator=g) * live[:, :, None]
    qr = qkv.clone().requires_grad_()
    q, k, v = [t.view(B, L, H, dh).permute(0, 2, 1, 3) for t in qr.split(d, dim=-1)]
    mask = O.sasrec_attention_mask(seq, causal, torch.float32)
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + mask, dim=-1)
    ctx_ref = (a @ v).permute(0, 2, 1, 3).reshape(B, L, d)
    ctx_ref.backward(dctx)
    ctx = torch.empty(B * L, d, device=DEV)
    lse = torch.empty(B, H, L, device=DEV)
    qd, sd = qkv.to(DEV).view(B * L, 3 * d), seq.to(DEV)
    ops.attn_fwd(qd, sd, H, dh, causal, ctx, lse)
    got = ctx.view(B, L, d).cpu()
    assert rel(got[live], ctx_ref[live]) < 5e-4
    assert float(got[~live].abs().max()) == 0.0 if (~live).any() else True      # padded query rows: zeros (dead values)
    dqkv = torch.full((B * L, 3 * d), 7.0, device=DEV)
    ops.attn_bwd(qd, sd, H, dh, causal, ctx, lse, dctx.to(DEV).view(B * L, d), dqkv)
    assert rel(dqkv.view(B, L, 3 * d), qr.grad) < 5e-4     # __expf-based softmax; bar is 1e-3


@pytest.mark.parametrize('loss_type', ['softmax', 'bpr'])
@pytest.mark.parametrize('d,N', [(32, 6), (64, 2), (128, 257), (128, 1025), (256, 40)])
@pytest.mark.parametrize('extras', [False, True])
def test_score_loss_fused(loss_type, d, N, extras):
    from unirec_b200 import ops
    g = gen(8)
    V, U, B = 500, 30, 7
    E = torch.randn(V, d, generator=g) * 0.3
    u = torch.randn(B, d, generator=g) * 0.3
    ids = torch.randint(0, V, (B, N), generator=g)
    label = torch.zeros(B, N, dtype=torch.int32)
    label[:, 0] = 1
    if extras and N > 3:
        label[2, 3] = 1                      # a second positive in one row (general label layout)
    ib = torch.randn(V, generator=g) * 0.1 if extras else None
    ubias = torch.randn(U, generator=g) * 0.1 if extras else None
    uid = torch.randint(1, U, (B,), generator=g)
    tau, clip = (0.7, 0.8) if extras else (1.0, -1.0)
    Er, ur = E.clone().requires_grad_(), u.clone().requires_grad_()
    s = O.inner_product_scores(ur, O.gather_rows(Er, ids))
    s = O.predict_layer(s, uid, ids, ubias, ib, tau, clip)
    s.retain_grad()
    loss_ref = O.cal_loss(s, label, loss_type)
    loss_ref.backward()
    scores, dscore = torch.empty(B, N, device=DEV), torch.empty(B, N, device=DEV)
    loss_vec, gu = torch.empty(B, device=DEV), torch.empty(B, d, device=DEV)
    n_pos = torch.zeros(1, device=DEV)
    lab_d = label.to(DEV)
    ops.count_positive(lab_d, n_pos)
    assert float(n_pos) == float(label.sum())
    ops.score_loss(E.to(DEV), u.to(DEV), ids.to(DEV), loss_type, label=lab_d if loss_type == 'softmax' else None,
                   item_bias=ib.to(DEV) if extras else None, user_bias=ubias.to(DEV) if extras else None, user_id=uid.to(DEV),
                   tau=tau, score_clip=clip, norm_dev=n_pos if loss_type == 'softmax' else None, norm_host=float(B * (N - 1)),
                   scores=scores, loss_vec=loss_vec, dscore=dscore, grad_user=gu)
    loss = torch.empty((), device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    if loss_type == 'softmax':
        ops.loss_finish(loss_vec, loss, denom_dev=n_pos, nan_flag=flag)
    else:
        ops.loss_finish(loss_vec, loss, denom_host=float(B), nan_flag=flag)
    assert rel(scores, s) < 1e-5
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    assert int(flag) == 0
    assert rel(gu, ur.grad) < 1e-4
    # dscore is dLoss/d(dot): compare through the dense table gradient it implies
    dE = torch.zeros(V, d, device=DEV)
    ops.scatter_add_rows(dE, ids.to(DEV), u.to(DEV), src_group=N, coef=dscore, coef_group=1, pad_id=-1)
    assert rel(dE, Er.grad) < 1e-4


def test_loss_finish_raises_nan_flag():
    from unirec_b200 import ops
    lv = torch.tensor([1.0, float('nan')], device=DEV)
    loss, flag = torch.empty((), device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.loss_finish(lv, loss, denom_host=2.0, nan_flag=flag)
    assert int(flag) == 1 and math.isnan(float(loss))


@pytest.mark.parametrize('mode', ['adam', 'adamw', 'sgd'])
def test_rowlist_optimizer_matches_lazy_adam(mode):
    from unirec_b200 import ops
    g = gen(9)
    V, d, B, N, L = 400, 64, 6, 5, 4
    table = torch.randn(V, d, generator=g)
    ids = torch.randint(0, V, (B, N), generator=g)
    ids[1, 1] = ids[0, 0]
    seq = torch.randint(0, V, (B, L), generator=g).to(torch.int32)
    u, coef = torch.randn(B, d, generator=g), torch.randn(B, N, generator=g)
    dX = torch.randn(B * L, d, generator=g)
    grad = torch.zeros(V, d)
    grad.index_add_(0, ids.reshape(-1), (coef.unsqueeze(-1) * u.unsqueeze(1)).reshape(-1, d))
    grad.index_add_(0, seq.reshape(-1).long(), dX)
    grad[0] = 0
    lr, wd = 0.01, (0.1 if mode != 'sgd' else 0.0)
    t_d = table.to(DEV)
    m_d, v_d = torch.zeros_like(t_d), torch.zeros_like(t_d)
    head = torch.full((V,), -1, dtype=torch.int32, device=DEV)
    n_ent = B * N + B * L
    nxt, uniq = torch.empty(n_ent, dtype=torch.int32, device=DEV), torch.empty(n_ent, dtype=torch.int32, device=DEV)
    n_uniq = torch.zeros(1, dtype=torch.int32, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    srcs = [(u.to(DEV), N, coef.to(DEV), 1, B * N), (dX.to(DEV), 1, None, 1, B * L)]
    ref_t, ref_m, ref_v = table.clone(), torch.zeros(V, d), torch.zeros(V, d)
    touched = torch.unique(torch.cat([ids.reshape(-1), seq.reshape(-1).long()]))
    touched = touched[touched > 0]
    for it in range(1, 3):
        n_uniq.zero_()
        ops.rowlist_link(head, ids.to(DEV), 0, nxt, uniq, n_uniq)
        ops.rowlist_link(head, seq.to(DEV), B * N, nxt, uniq, n_uniq)
        assert int(n_uniq) == touched.numel()
        sq = torch.zeros(1, device=DEV)
        ops.rowlist_apply(t_d, None, None, head, nxt, uniq, n_uniq, n_ent, srcs, 'sqnorm', sqnorm_out=sq)
        assert abs(float(sq) - float((grad.double() ** 2).sum())) <= 1e-4 * float((grad.double() ** 2).sum())
        ops.step_advance(step)
        ops.rowlist_apply(t_d, m_d, v_d, head, nxt, uniq, n_uniq, n_ent, srcs, mode, lr=lr, weight_decay=wd, step_dev=step)
        assert int((head != -1).sum()) == 0
        gt, pt = grad[touched], ref_t[touched]
        if mode == 'sgd':
            ref_t[touched] = pt - lr * gt
        else:
            if mode == 'adam':
                gt = gt + wd * pt
            else:
                pt = pt * (1 - lr * wd)
            ref_m[touched] = 0.9 * ref_m[touched] + 0.1 * gt
            ref_v[touched] = 0.999 * ref_v[touched] + 0.001 * gt * gt
            bc1, bc2 = 1 - 0.9 ** it, 1 - 0.999 ** it
            ref_t[touched] = pt - (lr / bc1) * ref_m[touched] / (ref_v[touched].sqrt() / math.sqrt(bc2) + 1e-8)
        assert rel(t_d, ref_t) < 1e-5
    # NaN skip: parameters untouched, lists still cleaned up
    before = t_d.clone()
    n_uniq.zero_()
    ops.rowlist_link(head, ids.to(DEV), 0, nxt, uniq, n_uniq)
    skip = torch.ones(1, dtype=torch.int32, device=DEV)
    ops.step_advance(step, skip)
    assert int(step) == 2
    ops.rowlist_apply(t_d, m_d, v_d, head, nxt, uniq, n_uniq, n_ent, srcs[:1], mode, lr=lr, step_dev=step, skip_flag=skip)
    assert torch.equal(before, t_d) and int((head != -1).sum()) == 0


def test_dense_opt_matches_torch_adam():
    from unirec_b200 import ops
    g = gen(10)
    n = 1000
    p0, grads = torch.randn(n, generator=g), [torch.randn(n, generator=g) for _ in range(3)]
    ref = p0.clone().requires_grad_()
    opt = torch.optim.Adam([ref], lr=0.01, weight_decay=0.05)
    p, m, v = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    for gr in grads:
        ref.grad = gr.clone()
        opt.step()
        ops.step_advance(step)
        ops.dense_opt(p, gr.to(DEV), m, v, 'adam', 0.01, weight_decay=0.05, step_dev=step)
    assert rel(p, ref) < 1e-5


def test_gru_gates_fwd_bwd():
    from unirec_b200 import ops
    g = gen(11)
    B, H = 5, 24
    gi, gh, hp = torch.randn(B, 3 * H, generator=g), torch.randn(B, 3 * H, generator=g), torch.randn(B, H, generator=g)
    dh = torch.randn(B, H, generator=g)
    gir, ghr, hpr = gi.clone().requires_grad_(), gh.clone().requires_grad_(), hp.clone().requires_grad_()
    r = torch.sigmoid(gir[:, :H] + ghr[:, :H])
    z = torch.sigmoid(gir[:, H:2 * H] + ghr[:, H:2 * H])
    n = torch.tanh(gir[:, 2 * H:] + r * ghr[:, 2 * H:])
    h = (1 - z) * n + z * hpr
    h.backward(dh)
    ho, save = torch.empty(B, H, device=DEV), torch.empty(B, 4 * H, device=DEV)
    ops.gru_gate_fwd(gi.to(DEV), 3 * H, gh.to(DEV), hp.to(DEV), ho, save, B, H)
    assert rel(ho, h) < 1e-5
    dgi, dgh, dhp = torch.empty(B, 3 * H, device=DEV), torch.empty(B, 3 * H, device=DEV), torch.empty(B, H, device=DEV)
    ops.gru_gate_bwd(dh.to(DEV), save, hp.to(DEV), dgi, 3 * H, dgh, dhp, B, H)
    assert rel(dgi, gir.grad) < 1e-4 and rel(dgh, ghr.grad) < 1e-4 and rel(dhp, hpr.grad) < 1e-4


# ---------------------------------------------------------------- tcgen05 tensor-core GEMM (TF32 in, fp32 accumulate)
TF32_TOL = 2e-3     # TF32 keeps 10 mantissa bits: ~5e-4 per product, bar for the tf32 mode is 1e-2 (bf16-class) / measured ~3e-4


@pytest.mark.parametrize('M,N,K', [(300, 128, 128), (1000, 384, 128), (777, 512, 512), (128, 256, 64), (5000, 128, 512)])
@pytest.mark.parametrize('act', [None, 'swish'])
def test_gemm_tc_nt(M, N, K, act):
    from unirec_b200 import ops
    assert ops.has_tensor_core_gemm()
    g = gen(21)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.1, torch.randn(N, generator=g)
    C, pre = torch.full((M, N), 7.0, device=DEV), torch.empty(M, N, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), C, M, N, K, transB=True, bias=b.to(DEV), act=act, preact=pre, precision=1)
    ref_pre = O.linear(A.double(), W.double(), b.double())
    ref = O.activation(ref_pre, act) if act else ref_pre
    assert rel(pre, ref_pre) < TF32_TOL and rel(C, ref) < TF32_TOL
    # accumulate (read-add-store epilogue)
    C2 = torch.ones(M, N, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), C2, M, N, K, transB=True, accumulate=True, precision=1)
    assert rel(C2, 1.0 + A.double() @ W.double().t()) < TF32_TOL


@pytest.mark.parametrize('Mo,No,T', [(128, 128, 5024), (384, 128, 51200), (512, 128, 4096), (128, 512, 4096)])
def test_gemm_tc_tn_weight_grad(Mo, No, T):
    """dW[Mo,No] += dY[T,Mo]^T @ X[T,No]: both operands MN-major, split-K over the token dimension."""
    from unirec_b200 import ops
    g = gen(22)
    dY, X = torch.randn(T, Mo, generator=g), torch.randn(T, No, generator=g)
    dW = torch.zeros(Mo, No, device=DEV)
    ops.gemm(dY.to(DEV), X.to(DEV), dW, Mo, No, T, transA=True, lda=Mo, accumulate=True, precision=1)
    ref = dY.double().t() @ X.double()
    assert rel(dW, ref) < TF32_TOL


@pytest.mark.parametrize('M,N,K', [(300, 128, 128), (1000, 128, 512), (5000, 512, 128), (51200, 128, 384)])
@pytest.mark.parametrize('precision', [1, 3])
def test_gemm_tc_nn_input_grad(M, N, K, precision):
    """dx[M,N] = dy[M,K] @ W[K,N]: A K-major, B MN-major (no transposed weight copy)."""
    from unirec_b200 import ops
    g = gen(24)
    dY, W = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g) * 0.1
    dX = torch.full((M, N), 3.0, device=DEV)
    ops.gemm(dY.to(DEV), W.to(DEV), dX, M, N, K, precision=precision)
    ref = dY.double() @ W.double()
    assert rel(dX, ref) < (TF32_TOL if precision == 1 else X3_TOL)
    ops.gemm(dY.to(DEV), W.to(DEV), dX, M, N, K, accumulate=True, precision=precision)
    assert rel(dX, 2 * ref) < (TF32_TOL if precision == 1 else X3_TOL)


X3_TOL = 1e-5       # 3xTF32 split: dropped terms are O(2^-22) per product -> fp32-class results


@pytest.mark.parametrize('M,N,K', [(300, 128, 128), (1000, 384, 128), (777, 512, 512), (51200, 128, 512)])
@pytest.mark.parametrize('act', [None, 'swish'])
def test_gemm_tc_3xtf32_nt(M, N, K, act):
    """precision=3: hi/lo operand split in shared memory, three tcgen05 MMAs per k-step -> same bar as the exact fp32 path."""
    from unirec_b200 import ops
    g = gen(25)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.1, torch.randn(N, generator=g)
    C, pre = torch.full((M, N), 7.0, device=DEV), torch.empty(M, N, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), C, M, N, K, transB=True, bias=b.to(DEV), act=act, preact=pre, precision=3)
    ref_pre = O.linear(A.double(), W.double(), b.double())
    ref = O.activation(ref_pre, act) if act else ref_pre
    assert rel(pre, ref_pre) < X3_TOL and rel(C, ref) < X3_TOL
    # not worse than the exact-FMA kernel against the float64 product
    C0 = torch.empty(M, N, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), C0, M, N, K, transB=True, bias=b.to(DEV), precision=0)
    assert rel(pre, ref_pre) < 8 * rel(C0, ref_pre) + 1e-6


@pytest.mark.parametrize('Mo,No,T', [(128, 128, 5024), (384, 128, 51200), (128, 512, 4096)])
def test_gemm_tc_3xtf32_tn_weight_grad(Mo, No, T):
    from unirec_b200 import ops
    g = gen(26)
    dY, X = torch.randn(T, Mo, generator=g), torch.randn(T, No, generator=g)
    dW = torch.zeros(Mo, No, device=DEV)
    ops.gemm(dY.to(DEV), X.to(DEV), dW, Mo, No, T, transA=True, lda=Mo, accumulate=True, precision=3)
    assert rel(dW, dY.double().t() @ X.double()) < X3_TOL


@pytest.mark.parametrize('M,N,K', [(4300, 128, 128), (5000, 384, 128), (51200, 128, 512)])      # (above the small-product SIMT route)
def test_gemm_tc_3xtf32_weight_lo_plane(M, N, K):
    """3xTF32 with the B operand's lo term fetched from a registered lo plane (ur_split_lo_f32 + ur_gemm_set_lo_plane) instead of
    the in-kernel split: NT (y = x W^T) and NN (dx = dy W) products of weights that live inside one parameter buffer; same bar, and
    the very same bits as the in-kernel split (both compute lo = tf32(w - trunc(w)))."""
    from unirec_b200 import ops
    g = gen(31)
    A, dY = torch.randn(M, K, generator=g).to(DEV), torch.randn(M, N, generator=g).to(DEV)
    flat = torch.zeros(1024 + N * K + 256, device=DEV)              # the weight sits in the middle of a larger buffer
    W = flat[1024:1024 + N * K].view(N, K)
    W.copy_(torch.randn(N, K, generator=g) * 0.1)
    C0, X0 = torch.empty(M, N, device=DEV), torch.empty(M, K, device=DEV)
    ops.gemm(A, W, C0, M, N, K, transB=True, precision=3)             # in-kernel split
    ops.gemm(dY, W, X0, M, K, N, precision=3)
    lo = torch.empty_like(flat)
    ops.split_lo(flat, lo)
    ops.gemm_set_lo_plane(flat, lo)
    try:
        C1, X1 = torch.empty(M, N, device=DEV), torch.empty(M, K, device=DEV)
        ops.gemm(A, W, C1, M, N, K, transB=True, precision=3)
        ops.gemm(dY, W, X1, M, K, N, precision=3)
        # a stale plane must show: the plane really is what the kernel reads
        lo.zero_()
        C2 = torch.empty(M, N, device=DEV)
        ops.gemm(A, W, C2, M, N, K, transB=True, precision=3)
    finally:
        ops.gemm_set_lo_plane(None, None)
    assert rel(C1, A.double() @ W.double().t()) < X3_TOL and rel(X1, dY.double() @ W.double()) < X3_TOL
    assert torch.equal(C1, C0) and torch.equal(X1, X0)
    assert not torch.equal(C2, C0) and rel(C2, A.double() @ W.double().t()) < TF32_TOL


@pytest.mark.parametrize('precision', [0, 1, 3])
@pytest.mark.parametrize('M,N,K', [(640, 512, 128), (5000, 128, 128)])
def test_gemm_fused_act_bwd_and_colsum(M, N, K, precision):
    """dh = (dz @ W) * act'(hpre) with colsum(dh) accumulated: the FFN backward epilogue fusion (modules.py:348-351 autograd)."""
    from unirec_b200 import ops
    g = gen(27)
    dZ, W, pre = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g) * 0.1, torch.randn(M, N, generator=g)
    dH = torch.empty(M, N, device=DEV)
    cs = torch.ones(N, device=DEV)
    ops.gemm_fused(dZ.to(DEV), W.to(DEV), dH, M, N, K, precision=precision, dact=pre.to(DEV), act='swish', colsum=cs)
    pre64 = pre.double().requires_grad_()
    O.activation(pre64, 'swish').sum().backward()
    ref = (dZ.double() @ W.double()) * pre64.grad
    tol = {0: 1e-5, 1: TF32_TOL, 3: X3_TOL}[precision]
    assert rel(dH, ref) < tol
    assert rel(cs, 1.0 + ref.sum(0)) < 10 * tol


def test_add_ln_bwd_fused_bias_grad():
    from unirec_b200 import ops
    g = gen(28)
    T, d = 1000, 128
    Z, dY, gamma = torch.randn(T, d, generator=g), torch.randn(T, d, generator=g), torch.rand(d, generator=g) + 0.5
    mean = Z.mean(1)
    rstd = 1.0 / torch.sqrt(Z.var(1, unbiased=False) + 1e-10)
    dZ, dg, db, dzs = (torch.empty(T, d, device=DEV), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV), torch.ones(d, device=DEV))
    ops.add_ln_bwd(Z.to(DEV), gamma.to(DEV), mean.to(DEV), rstd.to(DEV), dY.to(DEV), dZ, dg, db, dzsum=dzs)
    assert rel(dzs, 1.0 + dZ.double().sum(0).cpu()) < 1e-5


def test_transpose_small():
    from unirec_b200 import ops
    w = torch.randn(384, 128, generator=gen(23))
    out = torch.empty(128, 384, device=DEV)
    ops.transpose(w.to(DEV), out)
    assert torch.equal(out.cpu(), w.t().contiguous())


def test_pack_tokens_map():
    """Live positions = real items + position L-1; a sequence without real items keeps every position (csrc/pack.cu)."""
    from unirec_b200 import ops
    g = gen(31)
    B, L = 1500, 13          # > 1024 samples: exercises the multi-round block scan
    seq = torch.randint(1, 50, (B, L), generator=g, dtype=torch.int32)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    seq = torch.where(torch.arange(L)[None, :] >= (L - lens)[:, None], seq, torch.zeros_like(seq))
    seq[3, L - 1] = 0                      # pad at the output position
    seq[5, L - 3] = 0                      # pad between real items
    seq[7] = 0                             # empty history
    real = seq > 0
    keep = real.clone()
    keep[:, L - 1] = True
    keep |= ~real.any(1, keepdim=True)
    ref_src = torch.nonzero(keep.flatten()).flatten().to(torch.int32)
    n_ref = int(keep.sum())
    offs, src, inv = (torch.empty(B + 1, dtype=torch.int32, device=DEV), torch.full((B * L,), -7, dtype=torch.int32, device=DEV),
                      torch.empty(B * L, dtype=torch.int32, device=DEV))
    last, n = torch.empty(B, dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.pack_tokens(seq.to(DEV), offs, src, inv, last, n)
    assert int(n) == n_ref and int(offs[B]) == n_ref
    assert torch.equal(src[:n_ref].cpu(), ref_src)
    cnt = keep.sum(1)
    assert torch.equal(offs[:B].cpu().long(), torch.cumsum(cnt, 0) - cnt)
    assert torch.equal(last.cpu().long(), torch.cumsum(cnt, 0) - 1)
    inv_ref = torch.full((B * L,), -1, dtype=torch.int32)
    inv_ref[ref_src.long()] = torch.arange(n_ref, dtype=torch.int32)
    assert torch.equal(inv.cpu(), inv_ref)
    ops.pack_tokens(seq.to(DEV), offs, src, inv, last, n, keep_all=True)
    assert int(n) == B * L and torch.equal(src.cpu(), torch.arange(B * L, dtype=torch.int32))
    assert torch.equal(last.cpu().long(), torch.arange(B) * L + L - 1)


@pytest.mark.parametrize('H,B,L', [(768, 21, 6), (256, 37, 9), (64, 5, 3), (512, 16, 4)])
def test_persistent_gru_recurrence_matches_stepwise_path(H, B, L):
    """ur_gru_seq_fwd/bwd (one launch for all L steps; 16-row tiles for H <= 512, 8-row tiles above, e.g. the stock GRU.yaml
    hidden size 768) against the per-step GEMM + gate kernels and against torch's GRU cell arithmetic in float64."""
    from unirec_b200 import ops
    torch.manual_seed(H + B)
    k = 1.0 / H ** 0.5
    w_hh = (torch.rand(3 * H, H, device=DEV) * 2 - 1) * k
    b_hh = (torch.rand(3 * H, device=DEV) * 2 - 1) * k
    gi = torch.randn(B, L, 3 * H, device=DEV) * 0.5
    dh_last = torch.randn(B, H, device=DEV)
    # persistent
    hs = torch.zeros(L + 1, B, H, device=DEV)
    save = torch.empty(L, B, 4 * H, device=DEV)
    whh_t = ops.transpose(w_hh, torch.empty(H, 3 * H, device=DEV))
    ops.gru_seq_fwd(gi, whh_t, b_hh, hs, save, B, L, H)
    dgi = torch.empty(B, L, 3 * H, device=DEV)
    dgh = torch.empty(L, B, 3 * H, device=DEV)
    ops.gru_seq_bwd(dh_last, save, hs, w_hh, dgi, dgh, B, L, H)
    # float64 reference through autograd
    gi64 = gi.double().requires_grad_(True)
    w64, b64 = w_hh.double().requires_grad_(True), b_hh.double()
    h = torch.zeros(B, H, dtype=torch.float64, device=DEV)
    outs = []
    for t in range(L):
        gh = h @ w64.t() + b64
        g = gi64[:, t]
        r = torch.sigmoid(g[:, :H] + gh[:, :H])
        z = torch.sigmoid(g[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(g[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        outs.append(h)
    (h * dh_last.double()).sum().backward()
    ref_hs = torch.stack(outs)
    assert float((hs[1:].double() - ref_hs).abs().max()) < 2e-5
    assert float((dgi.double() - gi64.grad).abs().max()) < 2e-5 * max(1.0, float(gi64.grad.abs().max()))
    # dW_hh = sum_t dgh[t]^T h_{t-1} (formed by the GEMM in the engine): check the dgh the kernel emits through it
    dw = torch.einsum('tbj,tbk->jk', dgh.double(), hs[:-1].double())
    assert float((dw - w64.grad).abs().max()) < 5e-5 * max(1.0, float(w64.grad.abs().max()))


def test_gather_rows_bf16_is_exact_widening():
    from unirec_b200 import ops
    torch.manual_seed(2)
    table = (torch.randn(5000, 128, device=DEV) * 0.3).to(torch.bfloat16)
    for dt in (torch.int32, torch.int64):
        idx = torch.randint(0, 5000, (37, 11), device=DEV).to(dt)
        out = ops.gather_rows_bf16(table, idx)
        assert torch.equal(out, table[idx.long()].float())
