"""Micro-benchmark of the fused ResStack pair kernel (vfx_resstack_pair) at the vocoder's C = 64 shape (CUDA events).
usage: python tools/bench_pair.py [--B 32] [--dil 1,3,9,27,81,243,729,2187] [--iters 5]"""
import argparse, ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voicefixer_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--L", type=int, default=443646)
ap.add_argument("--dil", default="1,3,9,27,81,243,729,2187")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--impl", type=int, default=0, help="0 = pick, 1 = one CTA per tile (16-bit), 2 = two-CTA cluster pipeline, 3 = one CTA per tile, tf32 (TMEM residual stash)")
ap.add_argument("--C", type=int, default=64)
ap.add_argument("--prec", default="bf16")
args = ap.parse_args()
lib = _lib.load()
dev, C, B = "cuda:0", args.C, args.B
L = args.L if C == 64 else {128: 147882}[C] if args.L == 443646 else args.L
tf32 = args.prec == "tf32"
x = torch.randn(B, L, C, device=dev)
odt = torch.float32 if tf32 else torch.bfloat16
a = x if tf32 else torch.nn.functional.leaky_relu(x, 0.01).bfloat16()
a2 = torch.empty(B, L, C, device=dev, dtype=odt)
w1 = (torch.randn(3, C, C, device=dev) * 0.05).to(odt)
w2 = (torch.randn(3, C, C, device=dev) * 0.05).to(odt)
b1, b2 = torch.randn(C, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
scratch = torch.empty(int(lib.vfx_resstack_pair_scratch_bytes()), dtype=torch.uint8, device=dev)
for dil in [int(v) for v in args.dil.split(",")]:
    d = _lib.PairDesc()
    d.a, d.x, d.w1, d.b1, d.dilation, d.w2, d.b2 = a.data_ptr(), x.data_ptr(), w1.data_ptr(), b1.data_ptr(), dil, w2.data_ptr(), b2.data_ptr()
    d.B, d.L, d.C, d.write_raw, d.act, d.act_param = B, L, C, 1, _lib.ACT["lrelu"], 0.01
    d.precision, d.impl = _lib.PREC[args.prec], args.impl
    d.scratch, d.scratch_bytes = scratch.data_ptr(), scratch.numel()
    if tf32:
        d.x_out, d.stream_enc, d.stream_enc_out = a2.data_ptr(), 1, 1
    else:
        d.out_act = a2.data_ptr()
    _lib.check(lib.vfx_resstack_pair(ctypes.byref(d), st), "pair")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        _lib.check(lib.vfx_resstack_pair(ctypes.byref(d), st), "pair")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    els = B * L * C
    print(f"pair {args.prec} impl={args.impl} C={C} L={L} B={B} d={dil:5d}: {ms:7.3f} ms  {2 * 2 * 3 * C * els / ms / 1e9:7.1f} TF/s  "
          f"{12 * els / ms / 1e6:7.1f} GB/s (12 B/el traffic model)  {8 * els / ms / 1e6:7.1f} GB/s (8 B/el algorithmic)", flush=True)
