"""Live pin of the oracle against the UNMODIFIED reference on inputs that are NOT in the committed fixtures.

    python tests/golden/live_pin.py [seed]          # needs /root/reference; prints one JSON line

Same loading recipe as make_golden.py (stub-loader + seeded synthetic checkpoints under a temporary HOME, loaded by
the reference's own VoiceFixer() / Vocoder(44100)).  Cases: analysis at T in {2, 64, 128, 257} (pad/crop edges of the
64-frame UNet grid), Vocoder.forward at odd and even T, restore_inmem mode 0 on 1.3 s, restore_inmem mode 2
(train-mode BN + the dropout masks the reference actually drew, captured by forward hooks), and the
your_vocoder_func hook (base.py:126-129).  tests/test_oracle.py runs this in a subprocess when the reference is present."""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main(seed):
    os.environ["HOME"] = tempfile.mkdtemp(prefix="vfx_home_")
    import torch
    torch.set_num_threads(8)
    from voicefixer_b200 import synthetic
    from oracle import vf_oracle as O
    synthetic.write_checkpoints(os.environ["HOME"], seed=0)
    ana, voc = synthetic.make_analysis_state(0), synthetic.make_vocoder_state(1)
    import ref_loader
    ref_loader.install()
    from voicefixer.base import VoiceFixer as RefVoiceFixer          # unmodified reference
    from voicefixer.vocoder.base import Vocoder as RefVocoder
    ref, ref_voc = RefVoiceFixer(), RefVocoder(44100)
    model = ref._model

    def rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        assert a.shape == b.shape, (a.shape, b.shape)
        return float(np.sqrt(np.mean((a - b) ** 2)) / (np.sqrt(np.mean(b ** 2)) + 1e-30))

    rep = {}
    with torch.no_grad():
        for T in (2, 64, 128, 257):
            m = torch.rand(1, 1, T, 128, generator=torch.Generator().manual_seed(seed + T)) ** 4 * 30.0
            rep[f"analysis_T{T}"] = rel(O.analysis(m, ana), model(None, m)["mel"])
        for T in (5, 12):
            m = torch.rand(1, 1, T, 128, generator=torch.Generator().manual_seed(seed + 500 + T)) ** 4 * 30.0
            rep[f"vocoder_T{T}"] = rel(O.vocoder_forward(m, voc), ref_voc.forward(m, cuda=False))
        wav = synthetic.make_utterances(1, seconds=1.3, seed=seed + 1)[0]
        rep["restore_mode0_1.3s"] = rel(O.restore_inmem(wav, ana, voc, mode=0), ref.restore_inmem(wav, cuda=False, mode=0))

        # mode 2 end to end: capture the masks the reference's two Dropout(0.5) drew (x2 scaling: kept <=> out != 0)
        wav2 = synthetic.make_utterances(1, seconds=1.6, seed=seed + 2)[0]
        masks, hooks = [], []
        for mod in model.generator.denoiser:
            if isinstance(mod, torch.nn.Dropout):
                hooks.append(mod.register_forward_hook(lambda _m, inp, outp: masks.append((outp != 0) | (inp[0] == 0))))
        torch.manual_seed(seed + 3)
        out2 = ref.restore_inmem(wav2, cuda=False, mode=2)
        for h in hooks:
            h.remove()
        rep["restore_mode2_1.6s"] = rel(O.restore_inmem(wav2, ana, voc, mode=2, drop_masks_fn=lambda T: masks), out2)

        # the your_vocoder_func hook: the reference hands the callback a linear mel [1, 1, T, 128]
        seen = {}

        def my_vocoder(mel):
            seen["mel"] = mel.clone()
            return ref_voc.forward(mel, cuda=False) * 0.5
        ref2 = RefVoiceFixer()                                           # fresh module (mode 2 above moved BN running stats)
        out3 = ref2.restore_inmem(wav, cuda=False, mode=0, your_vocoder_func=my_vocoder)
        _, mel = O.frontend(torch.from_numpy(wav)[None], ana)
        o_mel = O.from_log(O.analysis(mel, ana))
        rep["hook_mel"] = rel(o_mel, seen["mel"])
        rep["hook_out"] = rel(O.trim_center(O.vocoder_forward(o_mel, voc) * 0.5, wav.shape[0]).squeeze(0).numpy(), out3)
    print(json.dumps(rep))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 9000)
