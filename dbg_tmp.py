import sys, ctypes as C; sys.path.insert(0,'.')
import torch, uavenv_b200
from target_allocation_ppo_transformer_b200 import _capi
L=_capi.load_policy()
L.uavpolicy_selftest_gemm_tile.argtypes=[C.c_void_p]*3+[C.c_int32]*2+[C.c_void_p]
torch.manual_seed(0)
for N,K in [(128,128),(256,128),(384,128),(128,256),(64,128)]:
    A=torch.randn(128,K,device="cuda").bfloat16(); W=(torch.randn(N,K,device="cuda")*0.2).bfloat16()
    D=torch.zeros(128,N,device="cuda")
    rc=L.uavpolicy_selftest_gemm_tile(A.data_ptr(),W.data_ptr(),D.data_ptr(),N,K,None)
    torch.cuda.synchronize()
    ref=A.float()@W.float().t()
    print(N,K,"rc",rc,"maxerr",float((D-ref).abs().max()),"refmax",float(ref.abs().max()))
