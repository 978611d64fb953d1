"""One profiled restore() step at the bench configuration; prints per-tag ms (profile level 2 = per dilation)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voicefixer_b200 import synthetic
from voicefixer_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng = Engine(synthetic.make_analysis_state(0), synthetic.make_vocoder_state(1), precision=os.environ.get("VFX_PRECISION", "bf16"))
wav = torch.from_numpy(synthetic.make_utterances(4, seconds=10.0, seed=1)).repeat(B // 4, 1).cuda()
out = torch.empty_like(wav)
for _ in range(2):
    eng.restore(wav, out=out)
torch.cuda.synchronize()
eng.set_option("profile", 2)
eng.restore(wav, out=out)
rep = eng.profile_report()
tot = sum(r["ms"] for r in rep.values())
print(f"total {tot:.2f} ms")
for t, r in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    tf = r["flops"] / r["ms"] / 1e9 if r["flops"] else 0
    print(f"{t:22s} n={r['count']:4d} {r['ms']:8.3f} ms  {tf:8.1f} TF/s")
