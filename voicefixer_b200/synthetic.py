"""Seeded synthetic checkpoints and inputs in the reference's layout.

The real Zenodo checkpoints cannot be fetched offline, so tests, smoke() and bench.py use
random weights of the reference architecture written in the reference checkpoint layout
(SURVEY 3.1):  ~/.cache/voicefixer/analysis_module/checkpoints/vf.ckpt  (flat state dict) and
~/.cache/voicefixer/synthesis_module/44100/model.ckpt-1490000_trimed.pt  ({"generator": sd}).
Values come from numpy's PCG64 so they are identical in the build container and on the GPU box.
"""
import math
import os
import numpy as np
import torch

ANALYSIS_CKPT = ".cache/voicefixer/analysis_module/checkpoints/vf.ckpt"
VOCODER_CKPT = ".cache/voicefixer/synthesis_module/44100/model.ckpt-1490000_trimed.pt"

UNET_ENC = [(2, 32), (32, 64), (64, 128), (128, 256), (256, 384), (384, 384)]
UNET_DEC = [(384, 384), (384, 384), (384, 256), (256, 128), (128, 64), (64, 32)]
VOC_UP = [(1024, 512, 7), (512, 256, 7), (256, 128, 3), (128, 64, 3)]


def _u(rng, shape, bound):
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def _bn(rng, sd, prefix, c):
    sd[prefix + ".weight"] = torch.from_numpy(rng.uniform(0.5, 1.5, c).astype(np.float32))
    sd[prefix + ".bias"] = torch.from_numpy((0.1 * rng.standard_normal(c)).astype(np.float32))
    sd[prefix + ".running_mean"] = torch.from_numpy((0.1 * rng.standard_normal(c)).astype(np.float32))
    sd[prefix + ".running_var"] = torch.from_numpy(rng.uniform(0.5, 1.5, c).astype(np.float32))
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def stft_conv_weights():
    """torchlibrosa 0.0.7 STFT conv kernels: real/imag of DFT matrix * periodic Hann,
    shape (1025, 1, 2048) fp32 (what f_helper.stft.conv_real/imag.weight hold in vf.ckpt)."""
    n = np.arange(2048)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / 2048)
    k = np.arange(1025)
    ang = -2 * np.pi * ((n[:, None] * k[None, :]) % 2048) / 2048
    real = (np.cos(ang) * win[:, None]).T[:, None, :]
    imag = (np.sin(ang) * win[:, None]).T[:, None, :]
    return torch.from_numpy(real.astype(np.float32)), torch.from_numpy(imag.astype(np.float32))


def htk_mel_fb():
    """mel.fb buffer as voicefixer/tools/mel_scale.py:173-238 builds it (fp32 torch, HTK,
    norm=None, f_min 0, f_max 22050, 1025 x 128)."""
    all_freqs = torch.linspace(0, 44100 // 2, 1025)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + 22050.0 / 700.0)
    m_pts = torch.linspace(m_min, m_max, 128 + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def _conv_block(rng, sd, p, cin, cout, gain):
    b1 = gain / math.sqrt(cin * 9)
    b2 = gain / math.sqrt(cout * 9)
    sd[p + ".conv1.weight"] = _u(rng, (cout, cin, 3, 3), b1)
    _bn(rng, sd, p + ".bn1", cin)
    sd[p + ".conv2.weight"] = _u(rng, (cout, cout, 3, 3), b2)
    _bn(rng, sd, p + ".bn2", cout)
    if cin != cout:
        bs = 1.0 / math.sqrt(cin)
        sd[p + ".shortcut.weight"] = _u(rng, (cout, cin, 1, 1), bs * 1.7)
        sd[p + ".shortcut.bias"] = _u(rng, (cout,), bs)


def make_analysis_state(seed=0, gain=1.7):
    """Flat state dict with vf.ckpt's key names (prefixes generator.*, f_helper.*, mel.fb)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    d = "generator.denoiser."
    for idx in ("0", "3", "7.bn", "8.bn", "9", "13"):
        _bn(rng, sd, d + idx, 1)
    # the first BN sees raw linear mel (values up to ~1e2): keep it O(1)
    sd[d + "0.running_var"] = torch.tensor([25.0])
    for idx, (o, i) in (("1", (256, 128)), ("4", (512, 256)), ("11", (512, 512)), ("15", (128, 512))):
        b = 1.0 / math.sqrt(i)
        sd[d + idx + ".weight"] = _u(rng, (o, i), b * gain)
        sd[d + idx + ".bias"] = _u(rng, (o,), b)
    for g in ("7", "8"):
        for layer in (0, 1):
            for suf in ("", "_reverse"):
                b = 1.0 / math.sqrt(256)
                sd[f"{d}{g}.gru.weight_ih_l{layer}{suf}"] = _u(rng, (768, 512), b)
                sd[f"{d}{g}.gru.weight_hh_l{layer}{suf}"] = _u(rng, (768, 256), b)
                sd[f"{d}{g}.gru.bias_ih_l{layer}{suf}"] = _u(rng, (768,), b)
                sd[f"{d}{g}.gru.bias_hh_l{layer}{suf}"] = _u(rng, (768,), b)
    u = "generator.unet."
    for i, (cin, cout) in enumerate(UNET_ENC, 1):
        for j in (1, 2, 3, 4):
            _conv_block(rng, sd, f"{u}encoder_block{i}.conv_block{j}", cin if j == 1 else cout, cout, gain)
    _conv_block(rng, sd, u + "conv_block7", 384, 384, gain)
    for i, (cin, cout) in enumerate(UNET_DEC, 1):
        p = f"{u}decoder_block{i}"
        sd[p + ".conv1.weight"] = _u(rng, (cin, cout, 3, 3), gain / math.sqrt(cin * 9 / 4))
        _bn(rng, sd, p + ".bn1", cin)
        for j in (2, 3, 4, 5):
            _conv_block(rng, sd, f"{p}.conv_block{j}", 2 * cout if j == 2 else cout, cout, gain)
    _conv_block(rng, sd, u + "after_conv_block1", 32, 32, gain)
    sd[u + "after_conv2.weight"] = _u(rng, (1, 32, 1, 1), 0.25 / math.sqrt(32))
    sd[u + "after_conv2.bias"] = _u(rng, (1,), 0.1)
    cr, ci = stft_conv_weights()
    sd["f_helper.stft.conv_real.weight"] = cr
    sd["f_helper.stft.conv_imag.weight"] = ci
    sd["mel.fb"] = htk_mel_fb()
    return sd


def _wn_conv(rng, sd, p, shape, fan_in, norm_dims, gain=1.0):
    b = 1.0 / math.sqrt(fan_in)
    v = _u(rng, shape, b * gain)
    nrm = v.pow(2).sum(norm_dims, keepdim=True).sqrt()
    g = nrm * torch.from_numpy(rng.uniform(0.7, 1.3, tuple(nrm.shape)).astype(np.float32))
    sd[p + ".parametrizations.weight.original0"] = g
    sd[p + ".parametrizations.weight.original1"] = v
    return b


def make_vocoder_state(seed=1, gain=1.4, old_style_keys=False):
    """State dict of the 44.1 kHz vocoder Generator (ckpt['generator'])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    cin = 128
    for i in (0, 2, 4, 6, 8):
        b = _wn_conv(rng, sd, f"condnet.{i}", (512, cin, 3), cin * 3, (1, 2), gain)
        sd[f"condnet.{i}.bias"] = _u(rng, (512,), b)
        cin = 512
    b = _wn_conv(rng, sd, "generator.1", (1024, 512, 7), 512 * 7, (1, 2), gain)
    sd["generator.1.bias"] = _u(rng, (1024,), b)
    for j, (ci, co, u) in enumerate(VOC_UP):
        p = f"generator.{3 + 3 * j}"
        # dead skip_conv: present in the checkpoint, never used
        sd[p + ".skip_conv.weight"] = _u(rng, (co, ci, 1), 1.0 / math.sqrt(ci))
        sd[p + ".skip_conv.bias"] = _u(rng, (co,), 1.0 / math.sqrt(ci))
        # ConvTranspose1d weight (Cin, Cout, 2u); weight-norm dim 0 = IN channels
        b = _wn_conv(rng, sd, p + ".layer", (ci, co, 2 * u), ci * 2, (1, 2), gain)
        sd[p + ".layer.bias"] = _u(rng, (co,), b)
        r = f"generator.{4 + 3 * j}"
        for i in range(8):
            for k in (1, 3):
                b = _wn_conv(rng, sd, f"{r}.layers.{i}.{k}", (co, co, 3), co * 3, (1, 2), gain)
                sd[f"{r}.layers.{i}.{k}.bias"] = _u(rng, (co,), b)
    b = _wn_conv(rng, sd, "generator.16", (1, 64, 7), 64 * 7, (1, 2), 0.5)
    sd["generator.16.bias"] = _u(rng, (1,), b)
    if old_style_keys:
        out = {}
        for k, v in sd.items():
            k = k.replace(".parametrizations.weight.original0", ".weight_g")
            k = k.replace(".parametrizations.weight.original1", ".weight_v")
            out[k] = v
        sd = out
    return sd


def write_checkpoints(home=None, seed=0, include_vocoder_in_vf=True):
    """Writes both checkpoint files under `home` (default: ~) in the reference layout."""
    home = home or os.path.expanduser("~")
    ana = make_analysis_state(seed)
    voc = make_vocoder_state(seed + 1)
    a_path = os.path.join(home, ANALYSIS_CKPT)
    v_path = os.path.join(home, VOCODER_CKPT)
    os.makedirs(os.path.dirname(a_path), exist_ok=True)
    os.makedirs(os.path.dirname(v_path), exist_ok=True)
    flat = dict(ana)
    if include_vocoder_in_vf:        # vf.ckpt also carries vocoder.model.* 
        for k, v in voc.items():
            flat["vocoder.model." + k] = v
    torch.save(flat, a_path)
    torch.save({"generator": voc}, v_path)
    return a_path, v_path


def make_utterances(batch, seconds=10.0, seed=1234):
    """Synthetic degraded speech-like utterances: harmonic stack with 3 Hz AM,
    one-pole low-pass, white noise, hard clip, peak-normalised to 0.9.  np float32 (B, L)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = int(round(seconds * 44100))
    t = np.arange(L, dtype=np.float64) / 44100.0
    out = np.empty((batch, L), dtype=np.float32)
    for b in range(batch):
        f0 = rng.uniform(90, 250)
        x = np.zeros(L)
        for h in range(1, 21):
            x += (0.3 / h) * np.sin(2 * np.pi * h * f0 * t + rng.uniform(0, 2 * np.pi))
        x *= 0.5 - 0.5 * np.cos(2 * np.pi * 3.0 * t + rng.uniform(0, 2 * np.pi))
        # crude low-pass: moving average of random width (2-8 kHz equivalent)
        w = int(44100 / rng.uniform(2000, 8000))
        x = np.convolve(x, np.ones(w) / w, mode="same")
        snr_db = rng.uniform(5, 30)
        x += rng.standard_normal(L) * np.sqrt(np.mean(x ** 2) / (10 ** (snr_db / 10)))
        x = np.clip(x, -rng.uniform(0.25, 1.0), rng.uniform(0.25, 1.0))
        out[b] = (0.9 * x / np.max(np.abs(x))).astype(np.float32)
    return out
