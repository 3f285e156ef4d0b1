"""ctypes binding of include/unirec_b200.h (the C ABI of the CUDA hot path).

The product path has NO fallback: if the shared library is missing or a symbol is absent, importing the ops
raises.  Build it with `python __graft_entry__.py build` (or `make -C unirec_b200/csrc`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libunirec_b200.so')

_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# name -> argument kinds; p pointer, i int, l int64, f float.  Mirrors include/unirec_b200.h one to one.
SIGNATURES = {
    'ur_version': '',
    'ur_has_tensor_core_gemm': '',
    'ur_gather_rows_f32': 'plipilpp',
    'ur_scatter_add_scalar_f32': 'ppillplp',
    'ur_gather_rows_bf16': 'plipilpp',
    'ur_scatter_add_rows_f32': 'plipilplpllp',
    'ur_pool_sum_fwd_f32': 'piplipfpppp' + 'ii' + 'p',
    'ur_seq_prep_ln_fwd_f32': 'ppppfpliippp' + 'pp' + 'pi' + 'pfi' + 'p',
    'ur_seq_prep_ln_bwd_f32': 'pppplii' + 'ppp' + 'pppp' + 'p' + 'pi' + 'pfi' + 'p',
    'ur_add_ln_fwd_f32': 'plplppflipl' + 'pp' + 'p' + 'pfipll' + 'p',
    'ur_add_ln_bwd_f32': 'plppp' + 'pl' + 'pl' + 'li' + 'pl' + 'ppp' + 'p' + 'pl' + 'pfipll' + 'p',
    'ur_dropout_rows_f32': 'pllipllppfi' + 'p',
    'ur_dropout_mask_f32': 'plipllpfii' + 'p',
    'ur_rng_advance': 'pp',
    'ur_gemm_f32': 'iilllplplplpipliip',
    'ur_gemm_simt_f32': 'iilllplplplpiplip',
    'ur_gemm_tc_f32': 'iilllplplplpipliip',
    'ur_gemm_fused_f32': 'iilllplplplpipliiplp' + 'pi' + 'p',
    'ur_gemm_simt_rows_f32': 'iilllplplplpipli' + 'pi' + 'p',
    'ur_pack_tokens': 'pliipppppp',
    'ur_zero_tail_rows_f32': 'plipl' + 'p',
    'ur_transpose_f32': 'pllpp',
    'ur_split_lo_f32': 'pplp',
    'ur_gemm_set_lo_plane': 'ppl',
    'ur_act_bwd_f32': 'pplip',
    'ur_colsum_accum_f32': 'plllp' + 'p' + 'p',
    'ur_attn_fwd_f32': 'ppliiiiipp' + 'ppp' + 'pfi' + 'p',
    'ur_attn_bwd_f32': 'ppliiiii' + 'pppp' + 'pppp' + 'pfi' + 'p',
    'ur_gru_gate_fwd_f32': 'plppppli' + 'p',
    'ur_gru_seq_fwd_f32': 'ppppplli' + 'p',
    'ur_gru_seq_bwd_f32': 'ppppppll' + 'i' + 'p',
    'ur_gru_gate_bwd_f32': 'pppplppli' + 'p',
    'ur_score_loss_fwd_bwd_f32': 'pipplipppp' + 'ffipf' + 'pppp' + 'p',
    'ur_count_positive_i32': 'plpp',
    'ur_loss_finish_f32': 'plpfppp',
    'ur_rowlist_link': 'ppillpppl' + 'iil' + 'p',
    'ur_rowlist_apply_f32': 'pppi' + 'pppp' + 'l' + 'plpl' + 'l' + 'plpl' + 'i' + 'fffff' + 'ppp' + 'p' + 'ppi' + 'pl' + 'p',
    'ur_ipc_export': 'ppp',
    'ur_ipc_open': 'plp',
    'ur_dense_opt_f32': 'ppppl' + 'i' + 'fffff' + 'ppp' + 'p',
    'ur_sqnorm_accum_f32': 'plpp',
    'ur_clip_coef_f32': 'pfpp',
    'ur_step_advance': 'ppp',
    'ur_build_batch': 'pplppp' + 'llpp' + 'iiii' + 'll' + 'pppp' + 'p',
    'ur_rank_target_f32': 'pipplpppfiipp',
    'ur_rank_count_f32': 'pliplpppppfiipp',
    'ur_rank_exclude_f32': 'piplpppppfiipplpp',
    'ur_shard_gather_rows_f32': 'pipiliipp',
    'ur_shard_localize': 'piliilpp',
    'ur_pack_ids_i32': 'pplipp',
    'ur_count_positive_packed': 'plpp',
    'ur_score_partial_f32': 'pippli' + 'ppp' + 'ffii' + 'ppp',
    'ur_score_merge_f32': 'pilifpppp' + 'p',
    'ur_score_dscore_f32': 'ppplii' + 'iffpp' + 'p',
    'ur_shard_scores_f32': 'pippli' + 'ppp' + 'fii' + 'pp',
    'ur_bpr_from_scores_f32': 'plifffppp' + 'p',
    'ur_shard_grad_user_f32': 'pippliiipp',
}

_KIND = {'p': _P, 'i': _I, 'l': _L, 'f': _F}
_lib = None


class NativeLibraryError(ImportError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raise loudly when the CUDA library is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            'unirec_b200: %s not found. The B200 path has no CPU/eager fallback; build the CUDA library with '
            '`python __graft_entry__.py build` or `make -C unirec_b200/csrc`.' % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, sig in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as e:
            raise NativeLibraryError('unirec_b200: symbol %s missing from %s (stale build?)' % (name, LIB_PATH)) from e
        fn.restype = _I
        fn.argtypes = [_KIND[c] for c in sig]
    _lib = handle
    return _lib


def exported_symbols():
    return sorted(SIGNATURES)


class NativeCallError(RuntimeError):
    pass


def check(rc, name):
    if rc != 0:
        if rc <= -1000:
            raise NativeCallError('%s: CUDA launch error %d' % (name, -rc - 1000))
        raise NativeCallError('%s: %s' % (name, {-1: 'bad argument', -2: 'unsupported shape'}.get(rc, 'error %d' % rc)))
