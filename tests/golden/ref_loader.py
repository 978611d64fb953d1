"""Stub-loader that imports the UNMODIFIED reference nn.Modules from /root/reference.

Test infrastructure only (used by make_golden.py in the build container; /root/reference
does not exist on the GPU box).  It registers empty namespace packages so the reference's
downloading `__init__`s are skipped, and fakes for the third-party modules that are absent
from this image (librosa, torchlibrosa, soundfile, matplotlib).  The torchlibrosa STFT fake
restates torchlibrosa 0.0.7's published algorithm (DFT-matrix conv1d, periodic Hann,
reflect pad), which is what voicefixer/tools/modules/fDomainHelper.py:23-31 constructs.
"""
import sys, types, os, math
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("VOICEFIXER_REFERENCE", "/root/reference")


def _ns(name, path=None):
    m = types.ModuleType(name)
    if path is not None:
        m.__path__ = [path]
    sys.modules[name] = m
    return m


class _STFT(nn.Module):
    """torchlibrosa.stft.STFT restated (n_fft=win_length, hann periodic, center, reflect)."""

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann",
                 center=True, pad_mode="reflect", freeze_parameters=True):
        super().__init__()
        assert window == "hann" and pad_mode in ("reflect", "constant")
        self.n_fft, self.hop_length = n_fft, hop_length
        self.center, self.pad_mode = center, pad_mode
        n = np.arange(n_fft)
        win = 0.5 - 0.5 * np.cos(2 * np.pi * n / n_fft)          # periodic hann (fftbins=True)
        out_channels = n_fft // 2 + 1
        x, y = np.meshgrid(np.arange(n_fft), np.arange(n_fft))
        W = np.power(np.exp(-2 * np.pi * 1j / n_fft), x * y)
        self.conv_real = nn.Conv1d(1, out_channels, n_fft, stride=hop_length, bias=False)
        self.conv_imag = nn.Conv1d(1, out_channels, n_fft, stride=hop_length, bias=False)
        self.conv_real.weight.data = torch.Tensor(
            np.real(W[:, 0:out_channels] * win[:, None]).T)[:, None, :]
        self.conv_imag.weight.data = torch.Tensor(
            np.imag(W[:, 0:out_channels] * win[:, None]).T)[:, None, :]
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, input):
        x = input[:, None, :]
        if self.center:
            x = F.pad(x, pad=(self.n_fft // 2, self.n_fft // 2), mode=self.pad_mode)
        real = self.conv_real(x)
        imag = self.conv_imag(x)
        real = real[:, None, :, :].transpose(2, 3)
        imag = imag[:, None, :, :].transpose(2, 3)
        return real, imag


class _ISTFT(nn.Module):
    """Constructed by the reference but never called on the restore() path."""

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        raise NotImplementedError("ISTFT is not on the restore() path")


def install():
    if "voicefixer" in sys.modules and getattr(sys.modules["voicefixer"], "_stubbed", False):
        return
    vf = os.path.join(REF, "voicefixer")
    _ns("voicefixer", vf)._stubbed = True
    _ns("voicefixer.vocoder", os.path.join(vf, "vocoder"))
    _ns("voicefixer.vocoder.model", os.path.join(vf, "vocoder", "model"))
    _ns("voicefixer.restorer", os.path.join(vf, "restorer"))
    _ns("voicefixer.tools", os.path.join(vf, "tools"))
    _ns("voicefixer.tools.modules", os.path.join(vf, "tools", "modules"))
    lib = _ns("librosa"); _ns("librosa.display"); lib.display = sys.modules["librosa.display"]
    _ns("librosa.filters"); lib.filters = sys.modules["librosa.filters"]
    mpl = _ns("matplotlib"); _ns("matplotlib.pyplot"); mpl.pyplot = sys.modules["matplotlib.pyplot"]
    mpl.cm = types.SimpleNamespace()
    sf = _ns("soundfile"); sf.write = lambda *a, **k: None
    tl = _ns("torchlibrosa"); st = _ns("torchlibrosa.stft"); tl.stft = st
    st.STFT, st.ISTFT = _STFT, _ISTFT
    st.magphase = lambda r, i: None
    # tools/wav.py imports these names at module import time
    pq = _ns("voicefixer.tools.modules.pqmf"); pq.PQMF = object


def ref_modules():
    """Returns the reference classes (imported from /root/reference, unmodified)."""
    install()
    from voicefixer.vocoder.config import Config
    Config.refresh(44100)
    from voicefixer.vocoder.model.generator import Generator as VocGenerator
    from voicefixer.restorer.model import Generator as AnaGenerator
    from voicefixer.tools.mel_scale import MelScale
    from voicefixer.tools.modules.fDomainHelper import FDomainHelper
    from voicefixer.vocoder.model import util as vutil
    from voicefixer.tools import pytorch_util as putil
    return dict(Config=Config, VocGenerator=VocGenerator, AnaGenerator=AnaGenerator,
                MelScale=MelScale, FDomainHelper=FDomainHelper, vutil=vutil, putil=putil)
