import sys, ctypes; sys.path.insert(0,'/root/repo')
import torch, tneq_b200
from tneq_b200 import _lib
lib=_lib.load()
torch.zeros(1,device='cuda')
for al in (0,1):
  for sk in (0,1):
    out=(ctypes.c_int*4)()
    rc=lib.tnq_gemm_kernel_attrs(al,sk,out); print(al,sk,rc,list(out))
