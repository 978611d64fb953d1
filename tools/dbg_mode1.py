import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import rel_rms
from voicefixer_b200 import synthetic
from voicefixer_b200.engine import Engine
from oracle import vf_oracle as O
ana, voc = synthetic.make_analysis_state(0), synthetic.make_vocoder_state(1)
eng = Engine(ana, voc, precision="fp32")
for seed in (43, 51, 52, 53):
    wav = synthetic.make_utterances(1, seconds=0.6, seed=seed)[0]
    x, cut = eng.hf_cut(wav[None])
    ref = O.remove_higher_frequency(wav)
    # oracle cut index
    n_fft, hop = 2048, 512
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)).astype(np.float32)
    xp = np.pad(wav, (1024, 1024)); nfr = 1 + (len(xp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(nfr)[:, None]
    S = np.abs(np.fft.rfft(xp[idx] * win[None, :], axis=1).T.astype(np.complex64))
    f = np.log10(S + 1e-8); f[f < 0] = 0; e = f.sum(1); thr = e.sum() * 0.95
    cur, i = e[0], 0
    while i < 1025 and cur < thr:
        cur += e[i + 1]; i += 1
    print(seed, "cut dev", int(cut[0]), "oracle", i, "hf rel", rel_rms(x[0].cpu().numpy(), ref), "margin", (cur - thr) / thr)
wav = synthetic.make_utterances(1, seconds=0.6, seed=51)[0]
x, cut = eng.hf_cut(wav[None])
xo = O.remove_higher_frequency(wav)
print("hf_cut rel", rel_rms(x[0].cpu().numpy(), xo), "len", x.shape, xo.shape)
y_dev_on_oracle_x = eng.restore(xo[None]).cpu().numpy()
y_dev = eng.restore(x).cpu().numpy()
y_ref = O.restore_inmem(wav, ana, voc, mode=1)
print("restore(dev x) vs ref", rel_rms(y_dev, y_ref), " restore(oracle x) vs ref", rel_rms(y_dev_on_oracle_x, y_ref))
_, mel_d = None, eng.frontend(x)
sp_o, mel_o = O.frontend(torch.from_numpy(xo)[None], ana)
print("mel rel", rel_rms(mel_d.cpu().numpy(), mel_o[:, 0].numpy()))
ml_d = eng.analysis(mel_d); ml_o = O.analysis(mel_o, ana)
print("logmel rel", rel_rms(ml_d.cpu().numpy(), ml_o[:, 0].numpy()), "max", float(ml_o.max()), float(ml_o.min()))
