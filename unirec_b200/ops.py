"""Thin torch-tensor wrappers over the C ABI (include/unirec_b200.h).

PyTorch is plumbing here: it owns device memory and the stream; every function below passes raw device
pointers to the hand-written sm_100a kernels, launched on torch's current CUDA stream.  There is no CPU or
eager fallback: a non-CUDA tensor raises.
"""
import torch

from . import _cabi

ACT_CODES = {None: 0, 'none': 0, 'swish': 1, 'gelu': 2, 'relu': 3, 'tanh': 4, 'sigmoid': 5}
LOSS_CODES = {'softmax': 0, 'bpr': 1}
OPT_CODES = {'adam': 0, 'adamw': 1, 'sgd': 2, 'sqnorm': 3}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('unirec_b200 ops need CUDA tensors (no CPU fallback exists); got device %s' % t.device)
    if dtype is not None and t.dtype != dtype:
        raise TypeError('expected %s, got %s' % (dtype, t.dtype))
    return t.data_ptr()


def _f32(t):
    return _ptr(t, torch.float32)


def _idx_bits(idx):
    if idx.dtype == torch.int32:
        return 32
    if idx.dtype == torch.int64:
        return 64
    raise TypeError('index tensor must be int32 or int64, got %s' % idx.dtype)


# measurement hooks (bench.py): number of C-ABI kernel launches; CUDA-event pairs around one named entry point
# (TIMED_OP) or around every entry point (PROFILE dict).  All off by default.
LAUNCH_COUNT = 0
TIMED_OP = None
TIMED_OPS = None      # bench.py: set of entry points bracketed by CUDA events -> TIMED_EVENTS gets (name, start, end)
TIMED_EVENTS = []
PROFILE = None
GEMM_LOG = None       # bench.py: list of {'flops', 'bytes', 'live'} per GEMM launch (live = bounded by the device-side token count)


def _call(name, *args):
    global LAUNCH_COUNT
    LAUNCH_COUNT += 1
    fn = getattr(_cabi.lib(), name)
    if PROFILE is not None or name == TIMED_OP or (TIMED_OPS is not None and name in TIMED_OPS):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args)
        b.record()
        if name == TIMED_OP:
            TIMED_EVENTS.append((a, b))
        elif TIMED_OPS is not None and name in TIMED_OPS:
            TIMED_EVENTS.append((name, a, b))
        if PROFILE is not None:
            PROFILE.setdefault(name, []).append((a, b))
    else:
        rc = fn(*args)
    _cabi.check(rc, name)


def version():
    return _cabi.lib().ur_version()


def has_tensor_core_gemm():
    return bool(_cabi.lib().ur_has_tensor_core_gemm())


# ------------------------------------------------------------------ rows
def gather_rows(table, idx, out=None):
    assert table.dim() == 2 and table.is_contiguous() and idx.is_contiguous()
    d = table.shape[1]
    if out is None:
        out = torch.empty(*idx.shape, d, dtype=torch.float32, device=table.device)
    _call('ur_gather_rows_f32', _f32(table), table.shape[0], d, _ptr(idx), _idx_bits(idx), idx.numel(), _f32(out), _stream())
    return out


def scatter_add_scalar(out, idx, src, idx_group=1, idx_mask=-1):
    """out[idx[e // idx_group] & idx_mask] += src.flat[e] (bias-vector gradients)."""
    assert out.is_contiguous() and idx.is_contiguous() and src.is_contiguous()
    _call('ur_scatter_add_scalar_f32', _f32(out), _ptr(idx), _idx_bits(idx), int(idx_group), int(idx_mask), _f32(src), src.numel(),
          _stream())
    return out


def gather_rows_bf16(table_bf16, idx, out=None):
    """Gather from a BF16 copy of a table, widened to fp32 (exact)."""
    assert table_bf16.dtype == torch.bfloat16 and table_bf16.is_contiguous() and idx.is_contiguous()
    d = table_bf16.shape[1]
    if out is None:
        out = torch.empty(*idx.shape, d, dtype=torch.float32, device=table_bf16.device)
    _call('ur_gather_rows_bf16', _ptr(table_bf16), table_bf16.shape[0], d, _ptr(idx), _idx_bits(idx), idx.numel(), _f32(out), _stream())
    return out


def scatter_add_rows(grad, idx, src, src_group=1, coef=None, coef_group=1, pad_id=0):
    assert grad.is_contiguous() and idx.is_contiguous() and src.is_contiguous()
    _call('ur_scatter_add_rows_f32', _f32(grad), grad.shape[0], grad.shape[1], _ptr(idx), _idx_bits(idx), idx.numel(),
          _f32(src), src_group, _f32(coef), coef_group, pad_id, _stream())
    return grad


def pool_sum_fwd(table, item_seq, item_seq_len, alpha, user_table=None, user_id=None, out=None, coeff_out=None, world=1, rank=0):
    B, L = item_seq.shape
    d = table.shape[1]
    if out is None:
        out = torch.empty(B, d, dtype=torch.float32, device=table.device)
    _call('ur_pool_sum_fwd_f32', _f32(table), d, _ptr(item_seq, torch.int32), B, L, _ptr(item_seq_len, torch.int64),
          float(alpha), _f32(user_table), _ptr(user_id, torch.int64) if user_id is not None else None, _f32(out),
          _f32(coeff_out), int(world), int(rank), _stream())
    return out


# ------------------------------------------------------------------ layer norm family
def _i32(t):
    return _ptr(t, torch.int32) if t is not None else None


class Drop:
    """One dropout site of one step: (rng device int64[2] = (seed, step), p, site id) plus the map from buffer rows to original
    token positions (row_pos tensor, or r * pos_mul + pos_add).  csrc/dropout.cuh."""
    __slots__ = ('rng', 'p', 'site', 'row_pos', 'pos_mul', 'pos_add')

    def __init__(self, rng, p, site, row_pos=None, pos_mul=1, pos_add=0):
        self.rng, self.p, self.site, self.row_pos, self.pos_mul, self.pos_add = rng, float(p), int(site), row_pos, int(pos_mul), int(pos_add)


def _drop3(drop):
    """(rng pointer, p, site) of a Drop, or the identity triple."""
    if drop is None or drop.p <= 0.0:
        return None, 0.0, 0
    return _ptr(drop.rng, torch.int64), drop.p, drop.site


def _drop6(drop):
    if drop is None or drop.p <= 0.0:
        return None, 0.0, 0, None, 1, 0
    return _ptr(drop.rng, torch.int64), drop.p, drop.site, _i32(drop.row_pos), drop.pos_mul, drop.pos_add


def seq_prep_ln_fwd(table, pos, gamma, beta, eps, item_seq, Y, mean, rstd, tok_src=None, n_tok=None, shards=None, drop=None):
    """shards = (int64 device tensor of W peer pointers, W): `item_seq` holds global ids of a row-sharded table."""
    B, L = item_seq.shape
    _call('ur_seq_prep_ln_fwd_f32', _f32(table), _f32(pos), _f32(gamma), _f32(beta), float(eps),
          _ptr(item_seq, torch.int32), B, L, table.shape[1], _f32(Y), _f32(mean), _f32(rstd), _i32(tok_src), _i32(n_tok),
          _ptr(shards[0], torch.int64) if shards else None, int(shards[1]) if shards else 0, *_drop3(drop), _stream())
    return Y


def seq_prep_ln_bwd(table, pos, gamma, item_seq, mean, rstd, dY, dX, dgamma, dbeta, dpos, tok_inv=None, shards=None, drop=None):
    B, L = item_seq.shape
    _call('ur_seq_prep_ln_bwd_f32', _f32(table), _f32(pos), _f32(gamma), _ptr(item_seq, torch.int32), B, L,
          table.shape[1], _f32(mean), _f32(rstd), _f32(dY), _f32(dX), _f32(dgamma), _f32(dbeta), _f32(dpos), _i32(tok_inv),
          _ptr(shards[0], torch.int64) if shards else None, int(shards[1]) if shards else 0, *_drop3(drop), _stream())
    return dX


def dropout_rows(X, drop, rows=None, d=None, ld=None, rows_dev=None):
    """X[r, :] *= mask in place (no-op when the site is off)."""
    if drop is None or drop.p <= 0.0:
        return X
    d = d or X.shape[-1]
    rows = rows if rows is not None else X.numel() // d
    _call('ur_dropout_rows_f32', _f32(X), ld or d, rows, d, _i32(drop.row_pos), drop.pos_mul, drop.pos_add, _i32(rows_dev),
          _ptr(drop.rng, torch.int64), drop.p, drop.site, _stream())
    return X


def dropout_mask(rng, p, site, rows, d, row_pos=None, pos_mul=1, pos_add=0, flat=False):
    """The multipliers (0 or 1/(1-p)) of a dropout site as a [rows, d] tensor (test hook for the CPU oracle)."""
    out = torch.empty(rows, d, dtype=torch.float32, device=rng.device)
    _call('ur_dropout_mask_f32', _f32(out), rows, d, _i32(row_pos), int(pos_mul), int(pos_add), _ptr(rng, torch.int64), float(p),
          int(site), int(flat), _stream())
    return out


def rng_advance(rng):
    _call('ur_rng_advance', _ptr(rng, torch.int64), _stream())


def ipc_export(t):
    """(64-byte CUDA IPC handle, byte offset) of a device tensor, for mapping it into the peer processes (csrc/p2p.cu)."""
    import ctypes
    h = ctypes.create_string_buffer(64)
    off = ctypes.c_int64(0)
    _cabi.check(_cabi.lib().ur_ipc_export(_ptr(t), ctypes.cast(h, ctypes.c_void_p), ctypes.cast(ctypes.byref(off), ctypes.c_void_p)),
                'ur_ipc_export')
    return bytes(h.raw), int(off.value)


def ipc_open(handle, offset):
    """Device pointer (int) of a peer's exported buffer."""
    import ctypes
    h = ctypes.create_string_buffer(handle, 64)
    out = ctypes.c_int64(0)
    _cabi.check(_cabi.lib().ur_ipc_open(ctypes.cast(h, ctypes.c_void_p), int(offset), ctypes.cast(ctypes.byref(out), ctypes.c_void_p)),
                'ur_ipc_open')
    return int(out.value)


def pack_tokens(item_seq, offs, tok_src, tok_inv, last_tok, n_tok, keep_all=False):
    """Live-token map of a left-padded [B, L] id matrix (csrc/pack.cu)."""
    B, L = item_seq.shape
    _call('ur_pack_tokens', _ptr(item_seq, torch.int32), B, L, int(keep_all), _i32(offs), _i32(tok_src), _i32(tok_inv),
          _i32(last_tok), _i32(n_tok), _stream())


def zero_tail_rows(X, n_dev, width=None, ld=None):
    _call('ur_zero_tail_rows_f32', _f32(X), ld or X.shape[-1], width or X.shape[-1], _i32(n_dev), X.shape[0], _stream())


def add_ln_fwd(X, R, gamma, beta, eps, Y, mean, rstd, rows=None, d=None, ldx=None, ldr=None, ldy=None, rows_dev=None, drop=None):
    """X <- dropout(X) + R; Y = LN(X)."""
    d = d or X.shape[-1]
    rows = rows if rows is not None else X.numel() // d
    _call('ur_add_ln_fwd_f32', _f32(X), ldx or d, _f32(R), ldr or d, _f32(gamma), _f32(beta), float(eps), rows, d,
          _f32(Y), ldy or d, _f32(mean), _f32(rstd), _i32(rows_dev), *_drop6(drop), _stream())
    return Y


def add_ln_bwd(Z, gamma, mean, rstd, dY, dZ, dgamma, dbeta, dExtra=None, rows=None, d=None, ldz=None, lddy=None,
               ldde=None, lddz=None, dzsum=None, rows_dev=None, drop=None, dZdrop=None):
    """dZ = LN'(Z)(dY + dExtra).  With an active dropout site, dZdrop receives dZ * mask (gradient of the linear branch)."""
    d = d or Z.shape[-1]
    rows = rows if rows is not None else Z.numel() // d
    on = drop is not None and drop.p > 0.0
    if on and dZdrop is None:
        raise ValueError('add_ln_bwd: an active dropout site needs the dZdrop output')
    _call('ur_add_ln_bwd_f32', _f32(Z), ldz or d, _f32(gamma), _f32(mean), _f32(rstd), _f32(dY), lddy or d,
          _f32(dExtra), ldde or d, rows, d, _f32(dZ), lddz or d, _f32(dgamma), _f32(dbeta), _f32(dzsum), _i32(rows_dev),
          _f32(dZdrop) if on else None, d, *_drop6(drop), _stream())
    return dZ


# ------------------------------------------------------------------ GEMM
def gemm(A, B, C, M, N, K, transA=False, transB=False, lda=None, ldb=None, ldc=None, bias=None, act=None,
         preact=None, ldp=None, accumulate=False, precision=0, rows_dev=None):
    """C[M,N] (+)= act(op(A)[M,K] @ op(B)[K,N] + bias).  Row-major; default leading dims = stored row length.
    rows_dev: device int32 live token count -- bounds M (row-parallel products) or K (transA: reductions over tokens)."""
    if rows_dev is not None:
        return gemm_fused(A, B, C, M, N, K, transA=transA, transB=transB, lda=lda, ldb=ldb, ldc=ldc, bias=bias, act=act,
                          preact=preact, ldp=ldp, accumulate=accumulate, precision=precision, rows_dev=rows_dev)
    if lda is None:
        lda = M if transA else K
    if ldb is None:
        ldb = K if transB else N
    if ldc is None:
        ldc = N
    if GEMM_LOG is not None:
        GEMM_LOG.append({'flops': 2.0 * M * N * K, 'bytes': 4.0 * (M * K + K * N + M * N * (2 if preact is not None else 1)), 'live': False})
    _call('ur_gemm_f32', int(transA), int(transB), M, N, K, _f32(A), lda, _f32(B), ldb, _f32(C), ldc, _f32(bias),
          ACT_CODES[act], _f32(preact), ldp or N, int(accumulate), int(precision), _stream())
    return C


def gemm_fused(A, B, C, M, N, K, transA=False, transB=False, lda=None, ldb=None, ldc=None, bias=None, act=None,
               preact=None, ldp=None, accumulate=False, precision=0, dact=None, ldd=None, colsum=None, rows_dev=None):
    """gemm() plus two fused epilogue stages: C = (A @ B) * act'(dact) (activation backward) and colsum += column sums of C."""
    if lda is None:
        lda = M if transA else K
    if ldb is None:
        ldb = K if transB else N
    if ldc is None:
        ldc = N
    if GEMM_LOG is not None:
        GEMM_LOG.append({'flops': 2.0 * M * N * K, 'live': rows_dev is not None,
                         'bytes': 4.0 * (M * K + K * N + M * N * (1 + (preact is not None) + (dact is not None) + bool(accumulate)))})
    _call('ur_gemm_fused_f32', int(transA), int(transB), M, N, K, _f32(A), lda, _f32(B), ldb, _f32(C), ldc, _f32(bias),
          ACT_CODES[act], _f32(preact), ldp or N, int(accumulate), int(precision), _f32(dact), ldd or N, _f32(colsum),
          _i32(rows_dev), (2 if transA else 1) if rows_dev is not None else 0, _stream())
    return C


def split_lo(x, lo):
    """lo[i] = tf32(x[i] - trunc_tf32(x[i])): the lo plane of a weight buffer for the 3xTF32 GEMMs (refresh after every weight change)."""
    n = x.numel()
    assert lo.numel() >= n and n % 4 == 0
    _call('ur_split_lo_f32', _f32(x), _f32(lo), n, _stream())
    return lo


def gemm_set_lo_plane(base, lo):
    """Register (weight buffer, lo plane): 3xTF32 GEMMs whose B operand lies inside `base` read B's lo term from `lo`.  None clears."""
    if base is None:
        _call('ur_gemm_set_lo_plane', None, None, 0)
    else:
        _call('ur_gemm_set_lo_plane', _f32(base), _f32(lo), base.numel())


def transpose(w, out):
    """out[c, r] = w[r, c] (small weight matrices)."""
    rows, cols = w.shape
    _call('ur_transpose_f32', _f32(w), rows, cols, _f32(out), _stream())
    return out


def act_bwd(dY, preact, act):
    _call('ur_act_bwd_f32', _f32(dY), _f32(preact), dY.numel(), ACT_CODES[act], _stream())
    return dY


def colsum_accum(X, M, N, out, ldx=None, rows_dev=None):
    _call('ur_colsum_accum_f32', _f32(X), ldx or N, M, N, _f32(out), _i32(rows_dev), _stream())
    return out


# ------------------------------------------------------------------ attention
def attn_fwd(qkv, item_seq, H, dh, causal, ctx, lse, q_only_last=False, offs=None, tok_src=None, q_last=None, drop=None):
    B, L = item_seq.shape
    _call('ur_attn_fwd_f32', _f32(qkv), _ptr(item_seq, torch.int32), B, L, H, dh, int(causal), int(q_only_last),
          _f32(ctx), _f32(lse), _i32(offs), _i32(tok_src), _f32(q_last), *_drop3(drop), _stream())
    return ctx


def attn_bwd(qkv, item_seq, H, dh, causal, ctx, lse, dctx, dqkv, q_only_last=False, offs=None, tok_src=None, q_last=None,
             dq_last=None, drop=None):
    B, L = item_seq.shape
    _call('ur_attn_bwd_f32', _f32(qkv), _ptr(item_seq, torch.int32), B, L, H, dh, int(causal), int(q_only_last),
          _f32(ctx), _f32(lse), _f32(dctx), _f32(dqkv), _i32(offs), _i32(tok_src), _f32(q_last), _f32(dq_last), *_drop3(drop),
          _stream())
    return dqkv


# ------------------------------------------------------------------ GRU pointwise
def gru_gate_fwd(gi, ld_gi, gh, h_prev, h_out, save, B, H):
    _call('ur_gru_gate_fwd_f32', _f32(gi), ld_gi, _f32(gh), _f32(h_prev), _f32(h_out), _f32(save), B, H, _stream())


def gru_gate_bwd(dh, save, h_prev, dgi, ld_dgi, dgh, dh_prev, B, H):
    _call('ur_gru_gate_bwd_f32', _f32(dh), _f32(save), _f32(h_prev), _f32(dgi), ld_dgi, _f32(dgh), _f32(dh_prev), B, H,
          _stream())


def gru_seq_fwd(gi, whh_t, b_hh, hs, save, B, L, H):
    _call('ur_gru_seq_fwd_f32', _f32(gi), _f32(whh_t), _f32(b_hh), _f32(hs), _f32(save), B, L, H, _stream())


def gru_seq_bwd(dh_last, save, hs, whh, dgi, dgh_all, B, L, H):
    _call('ur_gru_seq_bwd_f32', _f32(dh_last), _f32(save), _f32(hs), _f32(whh), _f32(dgi), _f32(dgh_all), B, L, H, _stream())


# ------------------------------------------------------------------ fused score + loss
def score_loss(table, user_emb, item_id, loss_type, label=None, item_bias=None, user_bias=None, user_id=None, tau=1.0,
               score_clip=-1.0, norm_dev=None, norm_host=1.0, scores=None, loss_vec=None, dscore=None, grad_user=None):
    B, N = item_id.shape
    _call('ur_score_loss_fwd_bwd_f32', _f32(table), table.shape[1], _f32(user_emb), _ptr(item_id, torch.int64), B, N,
          _ptr(label, torch.int32) if label is not None else None, _f32(item_bias), _f32(user_bias),
          _ptr(user_id, torch.int64) if user_id is not None else None, float(tau), float(score_clip),
          LOSS_CODES[loss_type], _f32(norm_dev), float(norm_host), _f32(scores), _f32(loss_vec), _f32(dscore),
          _f32(grad_user), _stream())


def count_positive(label, out):
    _call('ur_count_positive_i32', _ptr(label, torch.int32), label.numel(), _f32(out), _stream())
    return out


def loss_finish(loss_vec, loss_out, denom_dev=None, denom_host=1.0, nan_flag=None):
    _call('ur_loss_finish_f32', _f32(loss_vec), loss_vec.numel(), _f32(denom_dev), float(denom_host), _f32(loss_out),
          _ptr(nan_flag, torch.int32) if nan_flag is not None else None, _stream())
    return loss_out


# ------------------------------------------------------------------ row-sparse optimizer
def rowlist_link(head, keys, entry_offset, nxt, uniq, n_uniq, pad_id=0, world=1, rank=0, key_mask=-1):
    """world > 1: `keys` are global ids of a row-sharded table (only owned entries are linked, under local rows); key_mask strips
    flag bits of packed ids (ops.pack_ids)."""
    _call('ur_rowlist_link', _ptr(head, torch.int32), _ptr(keys), _idx_bits(keys), keys.numel(), entry_offset,
          _ptr(nxt, torch.int32), _ptr(uniq, torch.int32), _ptr(n_uniq, torch.int32), pad_id, int(world), int(rank), int(key_mask),
          _stream())


def rowlist_apply(table, mom, var, head, nxt, uniq, n_uniq, max_uniq, sources, mode, lr=0.0, beta1=0.9, beta2=0.999,
                  eps=1e-8, weight_decay=0.0, step_dev=None, grad_scale_dev=None, skip_flag=None, sqnorm_out=None,
                  u_begin=None, u_end=None, small_ctas=False):
    """sources: list of 1 or 2 tuples (src, src_group, coef_or_None, coef_group, n_entries[, (peer_ptrs int64 [W], rows_per_part)])."""
    s0 = sources[0]
    s1 = sources[1] if len(sources) > 1 else (None, 1, None, 1, 0)
    parts = s1[5] if len(s1) > 5 else None
    _call('ur_rowlist_apply_f32', _f32(table), _f32(mom), _f32(var), table.shape[1], _ptr(head, torch.int32),
          _ptr(nxt, torch.int32), _ptr(uniq, torch.int32), _ptr(n_uniq, torch.int32), max_uniq,
          _f32(s0[0]), s0[1], _f32(s0[2]), s0[3], s0[4], _f32(s1[0]), s1[1], _f32(s1[2]), s1[3],
          OPT_CODES[mode], float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
          _ptr(step_dev, torch.int32) if step_dev is not None else None, _f32(grad_scale_dev),
          _ptr(skip_flag, torch.int32) if skip_flag is not None else None, _f32(sqnorm_out),
          _ptr(u_begin, torch.int32) if u_begin is not None else None,
          _ptr(u_end, torch.int32) if u_end is not None else None, int(small_ctas),
          _ptr(parts[0], torch.int64) if parts else None, int(parts[1]) if parts else 0, _stream())


def dense_opt(param, grad, mom, var, mode, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step_dev=None,
              grad_scale_dev=None, skip_flag=None):
    _call('ur_dense_opt_f32', _f32(param), _f32(grad), _f32(mom), _f32(var), param.numel(), OPT_CODES[mode], float(lr),
          float(beta1), float(beta2), float(eps), float(weight_decay),
          _ptr(step_dev, torch.int32) if step_dev is not None else None, _f32(grad_scale_dev),
          _ptr(skip_flag, torch.int32) if skip_flag is not None else None, _stream())


def sqnorm_accum(grad, sqnorm):
    _call('ur_sqnorm_accum_f32', _f32(grad), grad.numel(), _f32(sqnorm), _stream())


def clip_coef(sqnorm, max_norm, coef):
    _call('ur_clip_coef_f32', _f32(sqnorm), float(max_norm), _f32(coef), _stream())


def step_advance(step, skip_flag=None):
    _call('ur_step_advance', _ptr(step, torch.int32), _ptr(skip_flag, torch.int32) if skip_flag is not None else None,
          _stream())


# ------------------------------------------------------------------ device-side batch builder
MASK_MODES = {None: 0, 'none': 0, 'unorder': 1, 'autoregressive': 2}


def build_batch(user_id, pos_item, n_users, n_items, K, L, hist_ptr=None, hist_items=None, hist_sorted=None, alias_prob=None,
                alias_idx=None, mask_mode='unorder', seq_last=0, seed=0, step=0):
    """Returns (item_id [B,1+K] i64, label [B,1+K] i32, item_seq [B,L] i32 | None, item_seq_len [B] i64 | None)."""
    B = user_id.shape[0]
    dev = user_id.device
    item_id = torch.empty(B, 1 + K, dtype=torch.int64, device=dev)
    label = torch.empty(B, 1 + K, dtype=torch.int32, device=dev)
    item_seq = torch.empty(B, L, dtype=torch.int32, device=dev) if L > 0 else None
    seq_len = torch.empty(B, dtype=torch.int64, device=dev) if L > 0 else None
    _call('ur_build_batch', _ptr(user_id, torch.int64), _ptr(pos_item, torch.int64), B,
          _ptr(hist_ptr, torch.int64) if hist_ptr is not None else None,
          _ptr(hist_items, torch.int32) if hist_items is not None else None,
          _ptr(hist_sorted, torch.int32) if hist_sorted is not None else None, n_users, n_items, _f32(alias_prob),
          _ptr(alias_idx, torch.int32) if alias_idx is not None else None, K, L, MASK_MODES.get(mask_mode, 0), int(seq_last),
          int(seed) & 0x7FFFFFFFFFFFFFFF, int(step), _ptr(item_id), _ptr(label), _ptr(item_seq) if item_seq is not None else None,
          _ptr(seq_len) if seq_len is not None else None, _stream())
    return item_id, label, item_seq, seq_len


# ------------------------------------------------------------------ row-sharded tables (multi-GPU)
def shard_gather_rows(table_local, idx, world, rank, out):
    _call('ur_shard_gather_rows_f32', _f32(table_local), table_local.shape[1], _ptr(idx), _idx_bits(idx), idx.numel(), world,
          rank, _f32(out), _stream())
    return out


def shard_localize(idx, world, rank, out, pad_id=0):
    _call('ur_shard_localize', _ptr(idx), _idx_bits(idx), idx.numel(), world, rank, pad_id, _ptr(out, torch.int32), _stream())
    return out


PACKED_ID_MASK = 0x7FFFFFFF


def pack_ids(item_id, label, out):
    """out[e] = id | (label > 0) << 31 (int32): the one id tensor the row-sharded step all-gathers."""
    B, N = item_id.shape
    _call('ur_pack_ids_i32', _ptr(item_id, torch.int64), _ptr(label, torch.int32) if label is not None else None, B, N,
          _ptr(out, torch.int32), _stream())
    return out


def count_positive_packed(ids, out):
    _call('ur_count_positive_packed', _ptr(ids, torch.int32), ids.numel(), _f32(out), _stream())
    return out


def score_partial(table_local, user_emb, ids, world, rank, z, state, item_bias=None, user_bias=None, user_id=None, tau=1.0,
                  score_clip=-1.0):
    S, N = ids.shape
    _call('ur_score_partial_f32', _f32(table_local), table_local.shape[1], _f32(user_emb), _ptr(ids, torch.int32), S, N,
          _f32(item_bias), _f32(user_bias), _ptr(user_id, torch.int64) if user_id is not None else None, float(tau),
          float(score_clip), world, rank, _f32(z), _f32(state), _stream())


def score_merge(states, world, B, d, tau, norm_dev, loss_vec, lse_ny, grad_user):
    _call('ur_score_merge_f32', _f32(states), int(world), int(B), int(d), float(tau), _f32(norm_dev), _f32(loss_vec), _f32(lse_ny),
          _f32(grad_user), _stream())


def score_dscore(z, ids, lse_ny, world, rank, tau, score_clip, norm_dev, dscore):
    S, N = ids.shape
    _call('ur_score_dscore_f32', _f32(z), _ptr(ids, torch.int32), _f32(lse_ny), S, N, world, rank, float(tau), float(score_clip),
          _f32(norm_dev), _f32(dscore), _stream())


def shard_scores(table_local, user_emb, ids, world, rank, z, item_bias=None, user_bias=None, user_id=None, tau=1.0):
    S, N = ids.shape
    _call('ur_shard_scores_f32', _f32(table_local), table_local.shape[1], _f32(user_emb), _ptr(ids, torch.int32), S, N,
          _f32(item_bias), _f32(user_bias), _ptr(user_id, torch.int64) if user_id is not None else None, float(tau), world, rank,
          _f32(z), _stream())


def bpr_from_scores(z, tau, score_clip, norm, loss_vec, dscore, scores=None):
    B, N = z.shape
    _call('ur_bpr_from_scores_f32', _f32(z), B, N, float(tau), float(score_clip), float(norm), _f32(loss_vec), _f32(dscore),
          _f32(scores), _stream())


def shard_grad_user(table_local, ids, dscore, world, rank, out):
    S, N = ids.shape
    _call('ur_shard_grad_user_f32', _f32(table_local), table_local.shape[1], _ptr(ids, torch.int32), _f32(dscore), S, N, world, rank,
          _f32(out), _stream())


# ------------------------------------------------------------------ one-vs-all ranking (evaluation, csrc/evalrank.cu)
def _i64(t):
    return _ptr(t, torch.int64) if t is not None else None


def rank_target(table_local, user_emb, target, tscore, item_bias=None, user_bias=None, user_id=None, tau=1.0, world=1, rank=0):
    _call('ur_rank_target_f32', _f32(table_local), table_local.shape[1], _f32(user_emb), _i64(target), target.numel(),
          _f32(item_bias), _f32(user_bias), _i64(user_id), float(tau), int(world), int(rank), _f32(tscore), _stream())
    return tscore


def rank_count(table_local, user_emb, target, tscore, counts, item_bias=None, user_bias=None, user_id=None, tau=1.0, world=1,
               rank=0):
    _call('ur_rank_count_f32', _f32(table_local), table_local.shape[0], table_local.shape[1], _f32(user_emb), target.numel(),
          _i64(target), _f32(tscore), _f32(item_bias), _f32(user_bias), _i64(user_id), float(tau), int(world), int(rank),
          _i32(counts), _stream())
    return counts


def rank_exclude(table_local, user_emb, target, tscore, counts, item_bias=None, user_bias=None, user_id=None, tau=1.0, world=1,
                 rank=0, hist_ptr=None, hist_sorted=None):
    _call('ur_rank_exclude_f32', _f32(table_local), table_local.shape[1], _f32(user_emb), target.numel(), _i64(target),
          _f32(tscore), _f32(item_bias), _f32(user_bias), _i64(user_id), float(tau), int(world), int(rank), _i64(hist_ptr),
          _i32(hist_sorted), (hist_ptr.numel() - 1) if hist_ptr is not None else 0, _i32(counts), _stream())
    return counts
