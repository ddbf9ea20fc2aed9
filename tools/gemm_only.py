import sys, ctypes; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, tneq_b200
from tneq_b200 import _lib
lib = _lib.load(); dev = torch.device('cuda:0')
M,N,K = 16384,8192,128
A = torch.randn(M,K,device=dev); B = torch.randn(N,K,device=dev); C = torch.empty(M,N,device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(4):
    _lib.check(lib.tnq_gemm_tf32x3(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(C.data_ptr()), M,N,K,K,K,N,1,0,0,0,0,st))
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
