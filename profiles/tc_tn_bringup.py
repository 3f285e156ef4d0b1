"""Bring-up probe for the MN-major (TN) tcgen05 path: sweep descriptor parameters, print the error of each."""
import ctypes, itertools, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unirec_b200 import _cabi, ops
lib = _cabi.lib()
lib.ur_gemm_tc_debug_set.argtypes = [ctypes.c_int] * 6
T, Mo, No = 256, 128, 128
g = torch.Generator().manual_seed(0)
dY, X = torch.randn(T, Mo, generator=g), torch.randn(T, No, generator=g)
ref = (dY.double().t() @ X.double())
dYd, Xd = dY.cuda(), X.cuda()
def run(lbo, sbo, kstep, major, layout, swz):
    lib.ur_gemm_tc_debug_set(lbo, sbo, kstep, major, layout, swz)
    dW = torch.zeros(Mo, No, device='cuda')
    rc = lib.ur_gemm_tc_f32(1, 0, Mo, No, T, dYd.data_ptr(), Mo, Xd.data_ptr(), No, dW.data_ptr(), No, None, 0, None, No, 1, 1,
                            torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = dW.double().cpu()
    err = float((out - ref).abs().max() / ref.abs().max())
    errT = float((out - ref.t()).abs().max() / ref.abs().max())
    return rc, err, errT, float(out.abs().max())
for lbo, sbo, kstep, layout, swz in itertools.product([4096, 512], [512, 1024, 4096], [1024], [1, 2], [4, 3, 5, 6]):
    try:
        print(lbo, sbo, kstep, layout, swz, run(lbo, sbo, kstep, 1, layout, swz), flush=True)
    except Exception as e:
        print(lbo, sbo, kstep, layout, swz, 'EXC', str(e)[:80], flush=True)
        break
