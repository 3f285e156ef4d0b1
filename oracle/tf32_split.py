"""CPU restatement of the 3xTF32 operand split of csrc/gemm_tc.cu (TEST INFRASTRUCTURE ONLY).

The tensor pipe's kind::tf32 reads an fp32 word and ignores its low 13 mantissa bits.  The product path therefore uses
    hi(x) = trunc_tf32(x)            (what the hardware sees when it is handed the raw fp32 word)
    lo(x) = tf32_rn(x - hi(x))       (the subtraction is exact in fp32; the rounding to 10 mantissa bits is
                                      "add 0x1000 to the bit pattern, let the pipe drop the low 13 bits")
and accumulates  a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  in fp32.  The dropped terms (a_lo*b_lo and the rounding of lo) are
O(2^-21) relative to |a||b|, which is why `gemm_precision: tf32x3` meets the fp32 parity bar of the reference's nn.Linear
(unirec/model/modules.py:254-356) while plain TF32 (2^-11) does not.

The same bit operations as `lo1` in gemm_tc.cu / `split_lo_kernel` (ur_split_lo_f32), on numpy uint32 views.
"""
import numpy as np

MASK = np.uint32(0xFFFFE000)


def trunc_tf32(x):
    """What a kind::tf32 MMA multiplies when it reads the fp32 word x."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    return (x.view(np.uint32) & MASK).view(np.float32)


def lo_bits(x):
    """The word ur_split_lo_f32 stores for x: bits(x - trunc(x)) + 0x1000 (the pipe's truncation completes the rounding)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    r = (x - trunc_tf32(x)).astype(np.float32)
    return (r.view(np.uint32) + np.uint32(0x1000)).view(np.float32)


def lo_tf32(x):
    """lo operand as the tensor pipe sees it."""
    return trunc_tf32(lo_bits(x))


def matmul_3xtf32(a, b):
    """a [M,K] @ b[K,N] with the three-term split, every product exact (tf32 x tf32 fits fp32... computed in float64 here) and the
    accumulation in float64: isolates the SPLIT error from the accumulation-order error of any particular kernel."""
    ah, al = trunc_tf32(a).astype(np.float64), lo_tf32(a).astype(np.float64)
    bh, bl = trunc_tf32(b).astype(np.float64), lo_tf32(b).astype(np.float64)
    return al @ bh + ah @ bl + ah @ bh


def matmul_tf32(a, b):
    return trunc_tf32(a).astype(np.float64) @ trunc_tf32(b).astype(np.float64)
