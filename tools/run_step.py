"""Runs N restore() steps at the bench configuration (no profiling hooks): for ncu launch lists."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voicefixer_b200 import synthetic
from voicefixer_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = Engine(synthetic.make_analysis_state(0), synthetic.make_vocoder_state(1), precision=os.environ.get("VFX_PRECISION", "bf16"))
wav = torch.from_numpy(synthetic.make_utterances(4, seconds=10.0, seed=1)).repeat(B // 4, 1).cuda()
out = torch.empty_like(wav)
c0 = eng.launch_count()
for i in range(n):
    eng.restore(wav, out=out)
    torch.cuda.synchronize()
    print("step", i, "launches so far", eng.launch_count() - c0, flush=True)
