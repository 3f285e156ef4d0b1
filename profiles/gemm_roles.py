"""Where the warp roles of the tcgen05 GEMM spend their cycles (UR_TC_PROF=1 role clocks, see gemm_tc.cu: g_tc_prof).
Usage: UR_TC_PROF=1 python profiles/gemm_roles.py [shape substring ...]"""
import ctypes, os, sys
os.environ.setdefault('UR_TC_PROF', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unirec_b200 import _cabi, ops
import importlib.util
spec = importlib.util.spec_from_file_location('gemm_shapes', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'gemm_shapes.py'))
lib = _cabi.lib()
lib.ur_gemm_tc_prof.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
NAMES = ['prod wait empty', 'mma wait ready', 'mma wait tmem', 'mma issue', 'split wait full', 'split wait lo_empty', 'split work',
         'epi wait tmem_full', 'epi work', 'cta lifetime', 'k-blocks', 'units', '  mma instructions', '  mma commits', '  split A -> TMEM', '  split B']

def main():
    T, d, I = 51200, 128, 512
    r = lambda *s: torch.randn(*s, device='cuda')
    x, wqkv, w1, hact, w2 = r(T, d), r(3 * d, d), r(I, d), r(T, I), r(d, I)
    qkv, ha, hp, z, b1 = torch.empty(T, 3 * d, device='cuda'), torch.empty(T, I, device='cuda'), torch.empty(T, I, device='cuda'), torch.empty(T, d, device='cuda'), r(I)
    shapes = [('fwd qkv NT 51200x384x128', lambda P: ops.gemm(x, wqkv, qkv, T, 3 * d, d, transB=True, precision=P)),
              ('fwd ffn1 NT 51200x512x128 +act+preact', lambda P: ops.gemm(x, w1, ha, T, I, d, transB=True, bias=b1, act='swish', preact=hp, precision=P)),
              ('fwd ffn2 NT 51200x128x512', lambda P: ops.gemm(hact, w2, z, T, d, I, transB=True, precision=P)),
              ('dW qkv TN 384x128x51200', lambda P: ops.gemm(qkv, x, torch.zeros(3 * d, d, device='cuda'), 3 * d, d, T, transA=True, lda=3 * d, accumulate=True, precision=P))]
    if len(sys.argv) > 1:
        shapes = [sh for sh in shapes if any(a in sh[0] for a in sys.argv[1:])]
    out = (ctypes.c_ulonglong * 16)()
    for name, fn in shapes:
        for P in (1, 3):
            for _ in range(3): fn(P)
            lib.ur_gemm_tc_prof(out, 1)
            n = 10
            for _ in range(n): fn(P)
            lib.ur_gemm_tc_prof(out, 1)
            v = [out[i] / n for i in range(16)]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): fn(P)
            e1.record(); torch.cuda.synchronize()
            print('%s precision %d: %.1f us per launch, back to back (warm L2)' % (name, P, e0.elapsed_time(e1) * 50))
            if not int(os.environ.get('UR_TC_PROF', '1')): continue
            ctas = 148
            life = v[9] / ctas
            print('%s precision %d: CTA lifetime %.1f us (%.0f cycles), %.1f k-blocks, %.1f units per CTA' % (name, P, life / 1965, life, v[10] / ctas, v[11] / ctas))
            for i in (0, 1, 2, 3, 12, 13, 4, 5, 6, 14, 15, 7, 8):
                print('    %-22s %6.1f%% of lifetime  (%.0f cycles / k-block)' % (NAMES[i], 100 * v[i] / v[9], v[i] / max(v[10], 1)))

if __name__ == '__main__':
    main()
