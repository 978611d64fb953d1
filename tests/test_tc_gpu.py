"""GPU tests of the tcgen05/TMA/TMEM conv-GEMM kernel (impl=1) through the C ABI.

bf16 x bf16 products are exact in fp32, so against an fp32 torch conv of the SAME bf16-rounded
operands only the accumulation order differs: tolerance 2e-5 relative RMS on the fp32 output and
bf16 rounding (4e-3) on the activated bf16 output."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F
from conftest import golden, rel_rms

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("C,L,dil,B", [(64, 128, 1, 1), (64, 700, 3, 2), (128, 1000, 27, 1), (256, 333, 243, 2),
                                       (512, 200, 2187, 1), (64, 5000, 81, 3)])
def test_tc_conv1d_dilated(C, L, dil, B):
    from gpu_util import conv_gemm
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = _rnd(B, C, L, seed=1).bfloat16()
    w = _rnd(C, C, 3, seed=2, scale=0.1).bfloat16()
    b = _rnd(C, seed=3)
    res = _rnd(B, L, C, seed=4)
    ref = F.conv1d(x.float(), w.float(), b, dilation=dil, padding=dil) + res.permute(0, 2, 1)
    a = x.permute(0, 2, 1).contiguous()[:, None]
    wp = w.permute(2, 0, 1).contiguous()
    raw, act = conv_gemm(a, wp, [(0, -dil), (0, 0), (0, dil)], bias=b, residual=res[:, None].contiguous(),
                         want_act=True, act="lrelu", act_param=0.01, precision="bf16", impl=1)
    assert rel_rms(raw[:, 0].permute(0, 2, 1).cpu(), ref.cpu()) < 2e-5
    assert rel_rms(act[:, 0].permute(0, 2, 1).float().cpu(), F.leaky_relu(ref, 0.01).cpu()) < 4e-3


@pytest.mark.parametrize("Cin,Cout,H,W", [(32, 32, 64, 127), (64, 64, 32, 63), (128, 64, 16, 31), (64, 128, 24, 15),
                                          (256, 256, 16, 7), (384, 384, 32, 3)])
def test_tc_conv2d_3x3(Cin, Cout, H, W):
    from gpu_util import conv_gemm
    torch.backends.cudnn.allow_tf32 = False
    x = _rnd(2, Cin, H, W, seed=5).bfloat16()
    w = _rnd(Cout, Cin, 3, 3, seed=6, scale=0.1).bfloat16()
    ref = F.conv2d(x.float(), w.float(), padding=1)
    a = x.permute(0, 2, 3, 1).contiguous()
    wp = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    taps = [(kh - 1, kw - 1) for kh in range(3) for kw in range(3)]
    raw, _ = conv_gemm(a, wp, taps, precision="bf16", impl=1)
    assert rel_rms(raw.permute(0, 3, 1, 2).cpu(), ref.cpu()) < 2e-5


@pytest.mark.parametrize("u,Cin,Cout,L", [(7, 128, 64, 300), (3, 64, 64, 1000)])
def test_tc_conv_transpose1d(u, Cin, Cout, L):
    from gpu_util import conv_gemm
    x = _rnd(2, Cin, L, seed=7).bfloat16()
    w = _rnd(Cin, Cout, 2 * u, seed=8, scale=0.1).bfloat16()
    b = _rnd(Cout, seed=9)
    pad = u // 2 + u % 2
    ref = F.conv_transpose1d(x.float(), w.float(), b, stride=u, padding=pad, output_padding=u % 2)
    a = x.permute(0, 2, 1).contiguous()[:, None]
    wp = w.permute(2, 1, 0).contiguous()
    mat, nA = Cout * Cin, u - pad
    out = torch.zeros(2, 1, u * L, Cout, device=DEV)
    conv_gemm(a, wp, [(0, 0), (0, -1)], N=nA * Cout, w_off=[pad * mat, (pad + u) * mat], bias=b, bias_mod=Cout,
              sw=u, rw=0, OW=u * L, out_ld=Cout, out_raw=out, precision="bf16", impl=1)
    conv_gemm(a, wp, [(0, 1), (0, 0)], N=(u - nA) * Cout, w_off=[0, u * mat], bias=b, bias_mod=Cout,
              sw=u, rw=nA, OW=u * L, out_ld=Cout, out_raw=out, precision="bf16", impl=1)
    assert rel_rms(out[:, 0].permute(0, 2, 1).cpu(), ref.cpu()) < 2e-5


def test_tc_conv_transpose2d_into_concat_buffer():
    from gpu_util import conv_gemm
    Cin, Cout, H, W = 128, 64, 16, 31
    x = _rnd(2, Cin, H, W, seed=10).bfloat16()
    w = _rnd(Cin, Cout, 3, 3, seed=11, scale=0.1).bfloat16()
    ref = F.conv_transpose2d(x.float(), w.float(), stride=2)[:, :, :-1, :]
    OH, OW = 2 * H, 2 * W + 1
    a = x.permute(0, 2, 3, 1).contiguous()
    wp = w.permute(2, 3, 1, 0).reshape(9, Cout, Cin).contiguous()
    out = torch.zeros(2, OH, OW, 2 * Cout, device=DEV)          # first half of a concat buffer
    for rh in range(2):
        for rw in range(2):
            taps, offs = [], []
            for kh in ([1] if rh else [0, 2]):
                for kw in ([1] if rw else [0, 2]):
                    taps.append((-1 if kh == 2 else 0, -1 if kw == 2 else 0))
                    offs.append((kh * 3 + kw) * Cout * Cin)
            conv_gemm(a, wp, taps, Hq=H, Wq=W + 1, N=Cout, w_off=offs, sh=2, rh=rh, sw=2, rw=rw, OH=OH, OW=OW,
                      out_ld=2 * Cout, out_raw=out, precision="bf16", impl=1)
    assert rel_rms(out[..., :Cout].permute(0, 3, 1, 2).cpu(), ref.cpu()) < 2e-5
    assert float(out[..., Cout:].abs().max()) == 0.0


def test_tc_wide_n_and_k7_valid_conv():
    """N = 1024 (4 N-tiles of 256), 7 taps, 'valid' window on a padded input (vocoder pre-conv)."""
    from gpu_util import conv_gemm
    Cin, Cout, L = 512, 1024, 150
    x = _rnd(1, Cin, L + 6, seed=12).bfloat16()
    w = _rnd(Cout, Cin, 7, seed=13, scale=0.05).bfloat16()
    ref = F.conv1d(x.float(), w.float())
    a = x.permute(0, 2, 1).contiguous()[:, None]
    wp = w.permute(2, 0, 1).contiguous()
    raw, _ = conv_gemm(a, wp, [(0, k) for k in range(7)], Wq=L, OW=L, precision="bf16", impl=1)
    assert rel_rms(raw[:, 0].permute(0, 2, 1).cpu(), ref.cpu()) < 2e-5
