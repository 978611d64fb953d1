import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.set_num_threads(8)
from voicefixer_b200 import synthetic
from voicefixer_b200.engine import Engine
from oracle import vf_oracle as O
step = sys.argv[1]
ana, voc = synthetic.make_analysis_state(0), synthetic.make_vocoder_state(1)
wav = synthetic.make_utterances(1, seconds=0.6, seed=51)[0]
t0 = time.time()
if step == "cpu2":
    y = O.restore_inmem(wav, ana, voc, mode=2); print("cpu mode2 ok", y.shape, time.time() - t0, flush=True)
else:
    eng = Engine(ana, voc, precision="fp32"); print("engine ok", flush=True)
    if step == "gpu2":
        y = eng.restore(wav[None], mode=2); torch.cuda.synchronize(); print("gpu mode2 ok", y.shape, time.time() - t0, flush=True)
    if step == "cond":
        from voicefixer_b200.api import oracle_conditions
        cond = oracle_conditions(wav); print("cond", cond.shape, flush=True)
        y = eng.vocoder_cond(torch.from_numpy(cond), scale=2.0 ** 15); torch.cuda.synchronize(); print("vocoder_cond ok", y.shape, time.time() - t0, flush=True)
