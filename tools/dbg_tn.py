import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pyrddlgym_jax_b200.core import engine
dev = torch.device("cuda:0")
M, N, K = 128, 256, 16
ws = torch.full((4 * M * N,), 7.0, device=dev)
A = torch.randn(K, M, device=dev); B = torch.randn(K, N, device=dev)
C = torch.full((M, N), 5.0, device=dev)
engine.tcgemm_tn(False, M, N, K, A, M, B, N, C, ws, 1)
torch.cuda.synchronize()
ref = A.t() @ B
print("C nonzero", int((C != 0).sum()), "C max", float(C.abs().max()), "ref max", float(ref.abs().max()))
print("ws[:8]", ws[:8].tolist(), "ws nonzero in first MN", int((ws[:M*N] != 0).sum()), "ws==7 count", int((ws[:M*N] == 7).sum()))
print("C[0,:6]", C[0, :6].tolist(), "ref[0,:6]", ref[0, :6].tolist())
# correlation with candidates
c = C.cpu().double().numpy(); r = ref.cpu().double().numpy()
print("max |C-ref|", np.abs(c - r).max(), "max |C - ref^..|")
