"""Runs the GRU recurrence kernel alone (for ncu): B sequences, T steps."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voicefixer_b200 import _lib
lib = _lib.load()
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = "cuda:0"
gi = torch.randn(B, T, 1536, device=dev) * 0.5
whh = torch.randn(2, 256, 768, device=dev) * 0.06
bhh = torch.randn(2, 768, device=dev) * 0.06
out = torch.empty(B, T, 512, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.vfx_gru_layer(p(gi), p(whh), p(bhh), B, T, p(out), None), "gru")
    e1.record(); torch.cuda.synchronize()
    print(f"gru B={B} T={T}: {e0.elapsed_time(e1):.3f} ms  = {e0.elapsed_time(e1)/T*1e3:.2f} us/step")
