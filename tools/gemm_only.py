import sys, ctypes; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, tneq_b200
from tneq_b200 import _lib
lib = _lib.load(); dev = torch.device('cuda:0')
M,N,K = 16384,8192,128
A = torch.randn(M,K,device=dev); B = torch.randn(N,K,device=dev); C = torch.empty(M,N,device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(4):
    _lib.check(lib.tnq_gemm_tf32x3(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(C.data_ptr()), M,N,K,K,K,N,1,0,0,0,0,st))
# the large-K variant (3-stage ring, one TMEM accumulator per stage): the K = 2 chi^2 contraction of the bond-64 sweep
M2,N2,K2 = 16384,128,8192
A2 = torch.randn(M2,K2,device=dev); B2 = torch.randn(N2,K2,device=dev); C2 = torch.empty(M2,N2,device=dev)
for _ in range(4):
    _lib.check(lib.tnq_gemm_tf32x3(ctypes.c_void_p(A2.data_ptr()), ctypes.c_void_p(B2.data_ptr()), ctypes.c_void_p(C2.data_ptr()), M2,N2,K2,K2,K2,N2,1,0,0,0,0,st))
torch.cuda.synchronize()
ref = (A2.double() @ B2.double().T)
print("large-K max rel err vs float64:", ((C2.double()-ref).abs().max()/ref.abs().max()).item())
for (m_,n_,k_,a_,b_,c_) in ((M,N,K,A,B,C),(M2,N2,K2,A2,B2,C2)):
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.check(lib.tnq_gemm_tf32x3(ctypes.c_void_p(a_.data_ptr()), ctypes.c_void_p(b_.data_ptr()), ctypes.c_void_p(c_.data_ptr()), m_,n_,k_,k_,k_,n_,1,0,0,0,0,st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    print(f"gemm {m_}x{n_}x{k_}: {ms:.3f} ms, {2*m_*n_*k_/ms/1e9:.1f} TFLOP/s fp32-equivalent ({3*2*m_*n_*k_/ms/1e9:.1f} issued TF32)")
x = torch.randn(256,64,4096,2,device=dev)
out = torch.empty(256,4096,64,2,device=dev)
from ctypes import c_int64, c_void_p
dims=[256,4096,64,2]; strides=[x.stride(0), x.stride(2), x.stride(1), 1]
for _ in range(3):
    _lib.check(lib.tnq_permute_f32(c_void_p(x.data_ptr()), c_void_p(out.data_ptr()), 4, (c_int64*4)(*dims), (c_int64*4)(*strides), 2, 0, st))
torch.cuda.synchronize()
assert torch.equal(out, x.permute(0,2,1,3).contiguous())
e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    _lib.check(lib.tnq_permute_f32(c_void_p(x.data_ptr()), c_void_p(out.data_ptr()), 4, (c_int64*4)(*dims), (c_int64*4)(*strides), 2, 0, st))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/5
print(f"permute (256,64,4096,c64)->(256,4096,64,c64): {ms:.3f} ms, {2*x.numel()*4/ms/1e6:.0f} GB/s (read+write)")
