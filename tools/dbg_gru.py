import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voicefixer_b200 import _lib
lib = _lib.load()
dev = "cuda:0"
for B, T in ((5, 37), (2, 37), (4, 37), (5, 4), (32, 64)):
    torch.manual_seed(0)
    gru = torch.nn.GRU(512, 256, num_layers=1, bidirectional=True, batch_first=True).to(dev)
    x = torch.randn(B, T, 512, device=dev)
    with torch.no_grad():
        ref, _ = gru(x)
        wih = torch.cat([gru.weight_ih_l0, gru.weight_ih_l0_reverse], 0)
        bih = torch.cat([gru.bias_ih_l0, gru.bias_ih_l0_reverse], 0)
        gi = (x @ wih.t() + bih).contiguous()
        whh_t = torch.stack([gru.weight_hh_l0.t(), gru.weight_hh_l0_reverse.t()], 0).contiguous()
        bhh = torch.stack([gru.bias_hh_l0, gru.bias_hh_l0_reverse], 0).contiguous()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    for rep in range(3):
        out = torch.zeros(B, T, 512, device=dev)
        _lib.check(lib.vfx_gru_layer(p(gi), p(whh_t), p(bhh), B, T, p(out), None), "gru")
        torch.cuda.synchronize()
        err = (out - ref).abs()
        per = err.reshape(B, T, 2, 256).amax(dim=(1, 3))
        print(f"B={B} T={T} rep{rep} max err {err.max().item():.3e}  per (b,dir): {[['%.1e' % v for v in row] for row in per.tolist()]}")
