"""Prints relative-RMS errors of the CUDA path vs the reference goldens for each precision."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden, rel_rms
from voicefixer_b200 import synthetic
from voicefixer_b200.engine import Engine

ana, voc = synthetic.make_analysis_state(0), synthetic.make_vocoder_state(1)
for prec, tc in (("fp32", 1), ("tf32", 0), ("tf32", 1), ("fp16", 0), ("fp16", 1), ("bf16", 0), ("bf16", 1)):
    eng = Engine(ana, voc, precision=prec)
    if prec != "fp32":
        eng.set_option("use_tc", tc)
    res = {}
    for T in (1, 63, 65, 130):
        g = golden(f"analysis_T{T}")
        res[f"ana{T}"] = rel_rms(eng.analysis(g["mel"][:, 0]).cpu().numpy(), g["out"][:, 0])
    for T in (3, 20):
        g = golden(f"vocoder_T{T}")
        o = eng.vocoder(g["mel"][:, 0]).cpu().numpy()
        res[f"voc{T}"] = rel_rms(o, g["out"][:, 0])
        res[f"voc{T}_mae"] = float(np.mean(np.abs(o - g["out"][:, 0])))
    g = golden("restore_mode0")
    o = eng.restore(g["wav"][None]).cpu().numpy()
    res["restore"] = rel_rms(o, g["out"]); res["restore_mae"] = float(np.mean(np.abs(o - g["out"])))
    print(prec, "tc" if tc else "simt", " ".join(f"{k}={v:.2e}" for k, v in res.items()), flush=True)
    del eng

# BASELINE item sizes against the oracle (needs ~20 s of CPU): 10 s utterance, every precision
from oracle import vf_oracle as O
wav = synthetic.make_utterances(1, seconds=10.0, seed=1234)[0]
ref = O.restore_inmem(wav, ana, voc, mode=0)
for prec in ("fp32", "tf32", "fp16", "bf16"):
    o = Engine(ana, voc, precision=prec).restore(wav[None]).cpu().numpy()
    print(f"10s {prec}: rel_rms={rel_rms(o, ref):.3e} mean_abs={float(np.mean(np.abs(o - ref))):.3e}", flush=True)
