"""Generate tests/golden/*.npz from the UNMODIFIED reference (/root/reference), and pin the
oracle restatement (oracle/vf_oracle.py) against it.  Run in the build container only:

    python tests/golden/make_golden.py

The reference is imported through tests/golden/ref_loader.py (namespace packages that skip the
downloading __init__s + fakes for absent third-party modules); synthetic seeded checkpoints
(voicefixer_b200/synthetic.py) are written in the reference layout under a temporary HOME so
`voicefixer.base.VoiceFixer()` and `voicefixer.vocoder.base.Vocoder(44100)` load them through
their own code.  Every fixture stores the seeds, the inputs and the reference's outputs.
"""
import os, sys, tempfile, time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

SEED = 0


def main():
    tmp_home = tempfile.mkdtemp(prefix="vfx_home_")
    os.environ["HOME"] = tmp_home
    import torch
    torch.set_num_threads(8)
    from voicefixer_b200 import synthetic
    from oracle import vf_oracle as O
    synthetic.write_checkpoints(tmp_home, seed=SEED)
    ana = synthetic.make_analysis_state(SEED)
    voc = synthetic.make_vocoder_state(SEED + 1)

    import ref_loader
    ref_loader.install()
    from voicefixer.base import VoiceFixer as RefVoiceFixer          # unmodified reference
    from voicefixer.vocoder.base import Vocoder as RefVocoder
    ref = RefVoiceFixer()
    ref_voc = RefVocoder(44100)
    model = ref._model

    def rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.sqrt(np.mean((a - b) ** 2)) / (np.sqrt(np.mean(b ** 2)) + 1e-30))

    report = {}
    with torch.no_grad():
        # ---- 1. front end (a5, a6): wav -> sp, mel
        wav = synthetic.make_utterances(2, seconds=0.25, seed=11)
        sp, _, _ = model.f_helper.wav_to_spectrogram_phase(torch.from_numpy(wav)[:, None, :])
        mel = model.mel(sp.permute(0, 1, 3, 2)).permute(0, 1, 3, 2)
        o_sp, o_mel = O.frontend(torch.from_numpy(wav), ana)
        report["frontend_mel"] = rel(o_mel, mel)
        assert report["frontend_mel"] < 1e-6
        np.savez_compressed(os.path.join(HERE, "frontend.npz"), wav=wav, mel=mel.numpy(),
                            sp_slice=sp.numpy()[:, :, :4], seed=SEED)

        # ---- 2. analysis stage (a7-a11): mel -> log-mel, several T to hit pad/crop edges
        for T in (1, 63, 65, 130):
            g = torch.Generator().manual_seed(100 + T)
            m = torch.rand(2, 1, T, 128, generator=g) ** 4 * 30.0
            out = model(None, m)["mel"]
            o = O.analysis(m, ana)
            report[f"analysis_T{T}"] = rel(o, out)
            assert report[f"analysis_T{T}"] < 2e-5, report
            np.savez_compressed(os.path.join(HERE, f"analysis_T{T}.npz"), mel=m.numpy(),
                                out=out.numpy(), seed=SEED)

        # ---- 3. mode 2 analysis stage (a19): train-mode BN, dropout masks captured by hooks
        T = 70
        g = torch.Generator().manual_seed(777)
        m = torch.rand(1, 1, T, 128, generator=g) ** 4 * 30.0
        masks = []
        hooks = []
        for mod in model.generator.denoiser:
            if isinstance(mod, torch.nn.Dropout):
                hooks.append(mod.register_forward_hook(
                    lambda _m, inp, outp: masks.append((outp != 0) | (inp[0] == 0))))
        import copy
        sd_before = copy.deepcopy(model.state_dict())
        model.train()
        torch.manual_seed(4242)
        out2 = model(None, m)["mel"]
        model.eval()
        for h in hooks:
            h.remove()
        model.load_state_dict(sd_before)          # undo the running-stat side effect
        o2 = O.analysis(m, ana, train=True, drop_masks=masks)
        report["analysis_mode2"] = rel(o2, out2)
        assert report["analysis_mode2"] < 1e-4, report
        np.savez_compressed(os.path.join(HERE, "analysis_mode2.npz"), mel=m.numpy(), out=out2.numpy(),
                            mask0=np.packbits(masks[0].numpy()), mask1=np.packbits(masks[1].numpy()),
                            seed=SEED)

        # ---- 4. vocoder (a13-a17): Vocoder.forward on linear mel
        for T in (3, 20):
            g = torch.Generator().manual_seed(200 + T)
            m = torch.rand(2, 1, T, 128, generator=g) ** 4 * 30.0
            out = ref_voc.forward(m, cuda=False)
            o = O.vocoder_forward(m, voc)
            report[f"vocoder_T{T}"] = rel(o, out)
            assert report[f"vocoder_T{T}"] < 2e-5, report
            np.savez_compressed(os.path.join(HERE, f"vocoder_T{T}.npz"), mel=m.numpy(),
                                out=out.numpy(), seed=SEED)

        # ---- 5. end-to-end restore_inmem modes 0 and 2(no dropout unavailable -> mode 0 only)
        wav = synthetic.make_utterances(1, seconds=0.5, seed=21)[0]
        out = ref.restore_inmem(wav, cuda=False, mode=0)
        o = O.restore_inmem(wav, ana, voc, mode=0)
        report["restore_mode0_0.5s"] = rel(o, out)
        assert out.shape == (1, wav.shape[0]) and report["restore_mode0_0.5s"] < 1e-4, report
        np.savez_compressed(os.path.join(HERE, "restore_mode0.npz"), wav=wav, out=out, seed=SEED)

        # ---- 6. segmentation (a2, D4): 30 s + 0.3 s -> two independent segments, concatenated
        wav = synthetic.make_utterances(1, seconds=30.3, seed=31)[0]
        t0 = time.time()
        out = ref.restore_inmem(wav, cuda=False, mode=0)
        print("reference 30.3 s restore on CPU: %.1f s" % (time.time() - t0))
        o = O.restore_inmem(wav, ana, voc, mode=0)
        report["restore_segmented"] = rel(o, out)
        assert out.shape == (1, wav.shape[0]) and report["restore_segmented"] < 1e-4, report
        sl = [slice(0, 4096), slice(O.SEG_LEN - 4096, O.SEG_LEN + 4096), slice(out.shape[1] - 4096, out.shape[1]),
              slice(600000, 604096)]
        np.savez_compressed(os.path.join(HERE, "restore_segmented.npz"), wav_seed=31, seconds=30.3,
                            slices=np.array([[s.start, s.stop] for s in sl]),
                            out_slices=np.concatenate([out[0, s] for s in sl]),
                            mean_abs=np.mean(np.abs(out)), rms=np.sqrt(np.mean(out ** 2)), seed=SEED)

    for k, v in report.items():
        print(f"oracle vs reference  {k:28s} rel-rms {v:.3e}")
    with open(os.path.join(HERE, "ORACLE_PIN.txt"), "w") as f:
        f.write("oracle/vf_oracle.py vs unmodified reference modules (rel. RMS), make_golden.py\n")
        for k, v in report.items():
            f.write(f"{k} {v:.3e}\n")


if __name__ == "__main__":
    main()
