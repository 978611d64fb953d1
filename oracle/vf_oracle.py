"""CPU oracle for the VoiceFixer restore()/Vocoder path.  TEST INFRASTRUCTURE ONLY.

A plain PyTorch-fp32 (CPU) restatement of the reference algorithm, function by function,
each citing the reference file:line it follows (paths relative to /root/reference).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module; the product path (`voicefixer_b200`) never does and fails loudly when its
CUDA extension is missing.

Parity pin: this restatement is checked against the UNMODIFIED reference nn.Modules imported
from /root/reference (tests/golden/make_golden.py, run in the build container) and against the
committed outputs of those modules (tests/golden/*.npz).  The reference's own FLAC golden files
(test/utterance/target/*.flac) need the Zenodo checkpoints, which are not available offline;
they pin only the length arithmetic here (tests/test_oracle.py::test_fixture_lengths).

All functions take flat state dicts in the reference checkpoint layout:
  ana: keys of vf.ckpt           ("generator.denoiser.*", "generator.unet.*", "mel.fb",
                                  "f_helper.stft.conv_real.weight", ...)
  voc: keys of ckpt["generator"] ("condnet.0.bias", "...parametrizations.weight.original0/1"
                                  or old-style "...weight_g/weight_v")
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

SR = 44100
N_FFT = 2048
HOP = 441
N_MEL = 128
SEG_LEN = SR * 30                      # voicefixer/base.py:116
UPSAMPLE_SCALES = (7, 7, 3, 3)         # voicefixer/vocoder/config.py:20
MEL_W_A, MEL_W_B = 18.8927416350036, 0.0269863588184314   # vocoder/config.py:301,310


# --------------------------------------------------------------------------- front end
def stft_mag(wav, ana):
    """wav (B, L) -> sp (B, 1, T, 1025).
    torchlibrosa.stft.STFT.forward as built at tools/modules/fDomainHelper.py:23-31 (DFT-matrix
    conv1d, stride 441, reflect pad 1024) and spectrogram_phase fDomainHelper.py:81-86 with
    eps=1e-8 (wav_to_spectrogram_phase :88)."""
    x = wav.float()[:, None, :]
    x = F.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    real = F.conv1d(x, ana["f_helper.stft.conv_real.weight"], stride=HOP)
    imag = F.conv1d(x, ana["f_helper.stft.conv_imag.weight"], stride=HOP)
    real = real[:, None].transpose(2, 3)
    imag = imag[:, None].transpose(2, 3)
    return torch.clamp(real ** 2 + imag ** 2, 1e-8, np.inf) ** 0.5


def mel_scale(sp, ana):
    """sp (B,1,T,1025) -> mel (B,1,T,128): tools/mel_scale.py:63-77 through base.py:83."""
    return torch.matmul(sp, ana["mel.fb"])


def frontend(wav, ana):
    """VoiceFixer._pre, base.py:78-85."""
    sp = stft_mag(wav, ana)
    return sp, mel_scale(sp, ana)


# --------------------------------------------------------------------------- denoiser
def _bn1(x, ana, prefix, train):
    """nn.BatchNorm2d(1) on (B,1,T,F); eval: running stats; train: biased batch stats over
    (N,H,W).  restorer/model.py:69-99 (BatchNorm2d(1) members)."""
    w, b = ana[prefix + ".weight"], ana[prefix + ".bias"]
    if train:
        return F.batch_norm(x, None, None, w, b, True, 0.0, 1e-5)
    return F.batch_norm(x, ana[prefix + ".running_mean"], ana[prefix + ".running_var"], w, b,
                        False, 0.0, 1e-5)


def gru_layer_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one torch.nn.GRU layer, batch_first.  x (B,T,I) -> (B,T,H).
    Gate order r,z,n; n = tanh(W_in x + b_in + r*(W_hn h + b_hn)); h' = (1-z)*n + z*h."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi = x @ w_ih.t() + b_ih
    h = x.new_zeros(B, H)
    out = x.new_zeros(B, T, H)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gh = h @ w_hh.t() + b_hh
        i_r, i_z, i_n = gi[:, t].chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1 - z) * n + z * h
        out[:, t] = h
    return out


def bn_gru(x, ana, prefix, train):
    """BN_GRU.forward restorer/model.py:56-62: BN2d(1) -> 2-layer bidirectional GRU(512->256)."""
    x = _bn1(x, ana, prefix + ".bn", train).squeeze(1)
    for layer in (0, 1):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            g = prefix + ".gru."
            outs.append(gru_layer_dir(x, ana[g + f"weight_ih_l{layer}{suffix}"],
                                      ana[g + f"weight_hh_l{layer}{suffix}"],
                                      ana[g + f"bias_ih_l{layer}{suffix}"],
                                      ana[g + f"bias_hh_l{layer}{suffix}"], rev))
        x = torch.cat(outs, -1)
    return x.unsqueeze(1)


def denoiser(mel, ana, train=False, drop_masks=None):
    """nn.Sequential at restorer/model.py:69-99.  drop_masks: in train mode (mode 2) the two
    Dropout(0.5) keep-masks (bool, shapes (B,1,T,512)); None -> no dropout (p forced to 0,
    SURVEY 8c hygiene note: torch RNG-stream parity is not attainable)."""
    p = "generator.denoiser."
    x = _bn1(mel, ana, p + "0", train)
    x = F.relu(F.linear(x, ana[p + "1.weight"], ana[p + "1.bias"]))
    x = _bn1(x, ana, p + "3", train)
    x = F.linear(x, ana[p + "4.weight"], ana[p + "4.bias"])
    if train and drop_masks is not None:
        x = x * drop_masks[0].to(x.dtype) * 2.0
    x = F.relu(x)
    x = bn_gru(x, ana, p + "7", train)
    x = bn_gru(x, ana, p + "8", train)
    x = F.relu(_bn1(x, ana, p + "9", train))
    x = F.linear(x, ana[p + "11.weight"], ana[p + "11.bias"])
    if train and drop_masks is not None:
        x = x * drop_masks[1].to(x.dtype) * 2.0
    x = F.relu(_bn1(x, ana, p + "13", train))
    x = F.linear(x, ana[p + "15.weight"], ana[p + "15.bias"])
    return torch.sigmoid(x)


# --------------------------------------------------------------------------- UNet
def _bn2d(x, ana, prefix, train):
    w, b = ana[prefix + ".weight"], ana[prefix + ".bias"]
    if train:
        return F.batch_norm(x, None, None, w, b, True, 0.0, 1e-5)
    return F.batch_norm(x, ana[prefix + ".running_mean"], ana[prefix + ".running_var"], w, b,
                        False, 0.0, 1e-5)


def conv_block_res(x, ana, p, train):
    """ConvBlockRes.forward restorer/modules.py:68-76 (pre-activation, lrelu 0.01)."""
    origin = x
    x = F.conv2d(F.leaky_relu(_bn2d(x, ana, p + ".bn1", train), 0.01), ana[p + ".conv1.weight"],
                 padding=1)
    x = F.conv2d(F.leaky_relu(_bn2d(x, ana, p + ".bn2", train), 0.01), ana[p + ".conv2.weight"],
                 padding=1)
    if (p + ".shortcut.weight") in ana:
        return F.conv2d(origin, ana[p + ".shortcut.weight"], ana[p + ".shortcut.bias"]) + x
    return origin + x


def encoder_block(x, ana, p, train):
    """EncoderBlockRes.forward restorer/modules.py:97-104."""
    for i in (1, 2, 3, 4):
        x = conv_block_res(x, ana, f"{p}.conv_block{i}", train)
    return F.avg_pool2d(x, (2, 2)), x


def decoder_block(x, skip, ana, p, train):
    """DecoderBlockRes.forward restorer/modules.py:149-157 (relu(bn1) -> ConvT 3x3 s2 ->
    prune last time row -> cat -> 4 ConvBlockRes)."""
    x = F.conv_transpose2d(F.relu(_bn2d(x, ana, p + ".bn1", train)), ana[p + ".conv1.weight"],
                           stride=2)
    x = x[:, :, 0:-1, :]
    x = torch.cat((x, skip), 1)
    for i in (2, 3, 4, 5):
        x = conv_block_res(x, ana, f"{p}.conv_block{i}", train)
    return x


def unet(x, ana, train=False):
    """UNetResComplex_100Mb.forward restorer/model_kqq_bn.py:130-181.  x (B,2,T,128)."""
    u = "generator.unet."
    origin_len = x.shape[2]
    pad_len = int(np.ceil(origin_len / 64)) * 64 - origin_len
    x = F.pad(x, (0, 0, 0, pad_len))
    x = x[..., 0:x.shape[-1] - 1]
    skips = []
    for i in range(1, 7):
        x, s = encoder_block(x, ana, f"{u}encoder_block{i}", train)
        skips.append(s)
    x = conv_block_res(x, ana, u + "conv_block7", train)
    for i in range(1, 7):
        x = decoder_block(x, skips[6 - i], ana, f"{u}decoder_block{i}", train)
    x = conv_block_res(x, ana, u + "after_conv_block1", train)
    x = F.conv2d(x, ana[u + "after_conv2.weight"], ana[u + "after_conv2.bias"])
    x = F.pad(x, (0, 1))
    return x[:, :, 0:origin_len, :]


def to_log(x):
    """tools/pytorch_util.py:18-22."""
    assert torch.sum(x < 0) == 0
    return torch.log10(torch.clip(x, min=1e-8))


def from_log(x):
    """tools/pytorch_util.py:25-27."""
    return 10 ** torch.clip(x, max=5)


def analysis(mel, ana, train=False, drop_masks=None):
    """restorer Generator.forward restorer/model.py:103-120 -> out['mel'] (log10 domain)."""
    clean = denoiser(mel.clone(), ana, train, drop_masks) * mel
    x = to_log(clean)
    unet_in = torch.cat([to_log(mel), x], 1)
    return unet(unet_in, ana, train) + x


# --------------------------------------------------------------------------- vocoder
def _wn(voc, prefix):
    """torch weight_norm (dim=0): w = v * g/||v|| over all dims but 0.
    Accepts the parametrizations layout and the old weight_g/weight_v layout."""
    if prefix + ".parametrizations.weight.original0" in voc:
        g, v = voc[prefix + ".parametrizations.weight.original0"], voc[prefix + ".parametrizations.weight.original1"]
    elif prefix + ".weight_g" in voc:
        g, v = voc[prefix + ".weight_g"], voc[prefix + ".weight_v"]
    else:
        return voc[prefix + ".weight"]
    return torch._weight_norm(v, g, 0)


def mel_weight():
    """Config.get_mel_weight_torch vocoder/config.py:296-316: a*exp(b*k), k = 1..128."""
    k = torch.linspace(1, N_MEL, N_MEL)
    return MEL_W_A * torch.exp(MEL_W_B * k)


def vocoder_normalize(mel):
    """Vocoder.forward vocoder/base.py:51-54 + tr_amp_to_db/tr_normalize/tr_pre
    vocoder/model/util.py:8-36,69-80.  mel (B,1,T,128) linear -> cond (B,128,T + T%2 + 4)."""
    assert mel.size()[-1] == 128
    mel = mel / mel_weight()[None, None, None, :].type_as(mel)
    min_level = torch.exp(torch.tensor(-100.0) / 20 * torch.log(torch.tensor(10.0)))
    S = 20 * torch.log10(torch.maximum(min_level, torch.abs(mel))) - 20.0
    S = torch.clip(8.0 * ((S + 115.0) / 115.0) - 4.0, -4.0, 4.0)
    c = S[:, 0].transpose(1, 2)
    pad_tail = c.size(-1) % 2 + 4
    return torch.cat([c, torch.zeros(c.size(0), N_MEL, pad_tail) - 4.0], -1)


def res_stack(x, voc, p):
    """ResStack.forward modules.py:592-595 over the Sequentials built at :550-576."""
    for i in range(8):
        d = 3 ** i
        h = F.conv1d(F.leaky_relu(x, 0.01), _wn(voc, f"{p}.layers.{i}.1"), voc[f"{p}.layers.{i}.1.bias"],
                     dilation=d, padding=d)
        h = F.conv1d(F.leaky_relu(h, 0.01), _wn(voc, f"{p}.layers.{i}.3"), voc[f"{p}.layers.{i}.3.bias"],
                     padding=1)
        x = x + h
    return x


def upsample_net(x, voc, p, u):
    """UpsampleNet.forward modules.py:501-517 with org=False, no_skip=True: x + sin x, then
    weight-normed ConvTranspose1d(k=2u, s=u, p=u//2+u%2, output_padding=u%2).  The skip_conv
    result is discarded by the reference and is not computed here."""
    x = x + torch.sin(x)
    return F.conv_transpose1d(x, _wn(voc, p + ".layer"), voc[p + ".layer.bias"], stride=u,
                              padding=u // 2 + u % 2, output_padding=u % 2)


def vocoder_generator(cond, voc):
    """vocoder Generator.forward generator.py:127-145 (condnet :33-54, generator :73-100)."""
    x = cond
    for i in (0, 2, 4, 6, 8):
        x = F.elu(F.conv1d(x, _wn(voc, f"condnet.{i}"), voc[f"condnet.{i}.bias"], padding=1))
    x = F.conv1d(F.pad(x, (3, 3), mode="reflect"), _wn(voc, "generator.1"), voc["generator.1.bias"])
    x = F.leaky_relu(x, 0.2)
    for j, u in enumerate(UPSAMPLE_SCALES):
        x = upsample_net(x, voc, f"generator.{3 + 3 * j}", u)
        x = res_stack(x, voc, f"generator.{4 + 3 * j}")
        x = F.leaky_relu(x, 0.2)
    x = F.conv1d(F.pad(x, (3, 3), mode="reflect"), _wn(voc, "generator.16"), voc["generator.16.bias"])
    return torch.tanh(x)


def vocoder_forward(mel, voc):
    """Vocoder.forward vocoder/base.py:42-56: (B,1,T,128) linear mel -> (B,1,(T+T%2+4)*441)."""
    return vocoder_generator(vocoder_normalize(mel), voc)


# --------------------------------------------------------------------------- restore
def trim_center(est, ref_len):
    """VoiceFixer._trim_center base.py:63-76 (est longer than ref on this path)."""
    n = est.shape[-1]
    diff = abs(n - ref_len)
    if n == ref_len:
        return est
    assert n > ref_len
    est = est[..., int(diff // 2): -int(diff // 2)]
    return est[..., :min(n, ref_len)]


def remove_higher_frequency(wav, ratio=0.95):
    """VoiceFixer.remove_higher_frequency base.py:87-104 (mode 1) with librosa 0.10.1
    (Dockerfile:9) defaults restated: stft(n_fft=2048, hop=512, hann periodic, center=True,
    pad_mode='constant'), istft(length=None -> 512*(frames-1), window-sumsquare normalised)."""
    wav = np.asarray(wav, dtype=np.float32)
    n_fft, hop = 2048, 512
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)).astype(np.float32)
    x = np.pad(wav, (n_fft // 2, n_fft // 2), mode="constant")
    nfr = 1 + (len(x) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(nfr)[:, None]
    stft = np.fft.rfft(x[idx] * win[None, :], axis=1).T.astype(np.complex64)     # (1025, T)
    real, img = np.real(stft), np.imag(stft)
    mag = (real ** 2 + img ** 2) ** 0.5
    cos, sin = real / (mag + 1e-8), img / (mag + 1e-8)
    spec = np.abs(stft)
    feature = np.log10(spec + 1e-8)
    feature[feature < 0] = 0
    energy_level = np.sum(feature, axis=1)
    threshold = np.sum(energy_level) * ratio
    curent_level, i = energy_level[0], 0
    while i < energy_level.shape[0] and curent_level < threshold:
        curent_level += energy_level[i + 1, ...]
        i += 1
    spec[i:, ...] = 0
    stft = spec * cos + 1j * spec * sin
    frames = np.fft.irfft(stft.T, n=n_fft, axis=1).astype(np.float32) * win[None, :]
    out_len = n_fft + hop * (nfr - 1)
    y = np.zeros(out_len, dtype=np.float32)
    wss = np.zeros(out_len, dtype=np.float32)
    wsq = win ** 2
    for t in range(nfr):
        y[t * hop:t * hop + n_fft] += frames[t]
        wss[t * hop:t * hop + n_fft] += wsq
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2: n_fft // 2 + hop * (nfr - 1)]


def restore_inmem(wav, ana, voc, mode=0, drop_masks_fn=None):
    """VoiceFixer.restore_inmem base.py:106-139: np (L,) -> np (1, N) float32.  30 s segments,
    independent, concatenated.  mode 2 = train-mode BN (+ dropout if masks given)."""
    wav = np.asarray(wav, dtype=np.float32)
    train = (mode == 2)
    res = []
    break_point = SEG_LEN
    while break_point < wav.shape[0] + SEG_LEN:
        segment = wav[break_point - SEG_LEN: break_point]
        if mode == 1:
            segment = remove_higher_frequency(segment)
        _, mel = frontend(torch.from_numpy(np.ascontiguousarray(segment))[None], ana)
        masks = drop_masks_fn(mel.shape[2]) if (train and drop_masks_fn) else None
        out_mel = analysis(mel, ana, train, masks)
        out = vocoder_forward(from_log(out_mel), voc)
        out = trim_center(out, segment.shape[-1])
        res.append(out)
        break_point += SEG_LEN
    return torch.cat(res, -1).squeeze(0).numpy()


# --------------------------------------------------------------------------- oracle() front end
def slaney_htk_mel_basis():
    """librosa.filters.mel(sr=44100, n_fft=2048, htk=True, n_mels=128, fmin=0, fmax=22050),
    norm='slaney' default, built in float64 then cast to float32 (vocoder/model/util.py:115-123)."""
    n_mels, fmax = 128, 22050.0
    fftfreqs = np.linspace(0, SR / 2.0, 1 + N_FFT // 2)
    mmax = 2595.0 * np.log10(1.0 + fmax / 700.0)
    mels = np.linspace(0.0, mmax, n_mels + 2)
    mel_f = 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + N_FFT // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def oracle_cond(wav):
    """Vocoder.oracle front end vocoder/base.py:61-73: peak-normalise, |librosa.stft|
    (n_fft 2048, hop 441, hann, center, pad_mode='constant' per librosa 0.10.1), Slaney mel,
    dB-20, normalise, pre() -> cond (1,128,T+T%2+4)."""
    wav = np.asarray(wav, dtype=np.float32)
    wav = wav / np.max(np.abs(wav))
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N_FFT) / N_FFT)).astype(np.float32)
    x = np.pad(wav, (N_FFT // 2, N_FFT // 2), mode="constant")
    nfr = 1 + (len(x) - N_FFT) // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(nfr)[:, None]
    stft = np.abs(np.fft.rfft(x[idx] * win[None, :], axis=1).T.astype(np.complex64))
    mel = np.dot(slaney_htk_mel_basis(), stft)
    min_level = np.exp(-100 / 20 * np.log(10))
    S = 20 * np.log10(np.maximum(min_level, np.abs(mel))) - 20
    S = np.clip(8.0 * ((S + 115.0) / 115.0) - 4.0, -4.0, 4.0)
    c = torch.FloatTensor(np.transpose(S, (1, 0))).unsqueeze(0).transpose(1, 2)
    pad_tail = c.size(-1) % 2 + 4
    return torch.cat([c, torch.zeros(1, N_MEL, pad_tail) - 4.0], -1)


def oracle_wave(wav, voc):
    """Vocoder.oracle vocoder/base.py:58-77 up to (not including) the file write:
    returns float wav_re * 2**15 as np (1, 1, S)."""
    return (vocoder_generator(oracle_cond(wav), voc) * 2 ** 15).numpy()
