"""One large NT GEMM (ncu target): python tools/gpu_one_gemm.py [M N K]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vqa_attention_networks_b200 import _lib
L = _lib.load()
M, N, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (8192, 8192, 8192)
A = torch.randn(M, K, device="cuda").bfloat16()
B = torch.randn(N, K, device="cuda").bfloat16()
C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(3):
    _lib.check(L.vqa_b200_gemm(p(A), 0, K, p(B), 0, K, p(C), 1, N, M, N, K, None, None, 1, 0, 0, 0, None, 0, None, st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    _lib.check(L.vqa_b200_gemm(p(A), 0, K, p(B), 0, K, p(C), 1, N, M, N, K, None, None, 1, 0, 0, 0, None, 0, None, st))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
L.vqa_b200_debug_set_counters(p(dbg))
_lib.check(L.vqa_b200_gemm(p(A), 0, K, p(B), 0, K, p(C), 1, N, M, N, K, None, None, 1, 0, 0, 0, None, 0, None, st))
torch.cuda.synchronize()
L.vqa_b200_debug_set_counters(None)
d = dbg.tolist()
print("cta0 producer: empty-wait %d of %d cycles | cta1 producer: empty-wait %d of %d | MMA: full-wait %d, acc-wait %d of %d cycles" % (d[0], d[1], d[8], d[9], d[2], d[3], d[4]))
print("gemm %dx%dx%d: %.3f ms %.1f TFLOP/s (CTA2=%s)" % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9, os.environ.get("VQA_B200_CTA2")))
