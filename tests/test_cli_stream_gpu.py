"""GPU tests of the host-side callers of the hot path: the CLI `python -m voicefixer_b200` (mirror of
voicefixer/__main__.py:13-219) run for real on a folder of wav files, the batch and the streaming entry points."""
import os
import numpy as np
import pytest
from conftest import rel_rms

pytestmark = pytest.mark.gpu


@pytest.fixture()
def home(tmp_path, monkeypatch):
    from voicefixer_b200 import synthetic
    monkeypatch.setenv("HOME", str(tmp_path))
    synthetic.write_checkpoints(str(tmp_path), seed=0)
    return tmp_path


@pytest.mark.timeout(600)
def test_cli_folder_and_all_modes(home, states, monkeypatch, capsys):
    """main(): folder mode over three wav files, then `--mode all` on one file (three outputs named <name>-mode<k>.flac);
    outputs are compared with the CPU oracle at the reference's own acceptance bar (mean-abs < 0.01, test/test.py:35)."""
    from voicefixer_b200 import synthetic, wavio
    from voicefixer_b200.__main__ import main
    from oracle import vf_oracle as O
    monkeypatch.setenv("VFX_PRECISION", "tf32")
    src, dst = home / "in", home / "out"
    os.makedirs(src)
    wavs = synthetic.make_utterances(3, seconds=1.5, seed=61)
    for i, w in enumerate(wavs):
        wavio.save_wave(w[None], str(src / f"u{i}.wav"))
    (src / "notes.txt").write_text("ignored: not a .wav")
    assert main(["--infolder", str(src), "--outfolder", str(dst)]) == 0
    assert sorted(os.listdir(dst)) == ["u0.wav", "u1.wav", "u2.wav"]
    for i in range(3):
        x16 = wavio.load_mono(str(src / f"u{i}.wav"))                   # what the CLI read (int16 round trip)
        got = wavio.load_mono(str(dst / f"u{i}.wav"))
        ref = O.restore_inmem(x16, states[0], states[1], mode=0)[0]
        assert got.shape == ref.shape and float(np.mean(np.abs(got - ref))) < 1e-3
    assert "Processing" in capsys.readouterr().out
    assert main(["--infile", str(src / "u0.wav"), "--outfile", str(home / "r.flac"), "--mode", "all", "--silent"]) == 0
    n = wavs.shape[1]
    for mode, want in ((0, n), (1, 512 * (n // 512)), (2, n)):
        y = wavio.load_mono(str(home / f"r-mode{mode}.flac"))
        assert y.shape[0] == want and np.isfinite(y).all()
    with pytest.raises(ValueError, match="only support the .wav format"):
        main(["--infile", str(home / "r-mode0.flac"), "--outfile", str(home / "x.wav")])


@pytest.mark.timeout(600)
def test_restore_batch_equals_restore_inmem(home):
    """The batch entry point (one CUDA-graph replay for B utterances) returns, row by row, what the reference-shaped call
    restore_inmem returns for that utterance alone; pinned and pageable buffers take the same path."""
    from voicefixer_b200 import synthetic, api
    vf = api.VoiceFixer(precision="tf32")
    wavs = synthetic.make_utterances(3, seconds=1.2, seed=71)
    out = np.array(vf.restore_batch(wavs))
    for b in range(3):
        assert rel_rms(out[b], vf.restore_inmem(wavs[b], cuda=True, mode=0)[0]) < 1e-6
    pin_in, pin_out = vf.pinned_empty(wavs.shape), vf.pinned_empty(wavs.shape)
    pin_in[...] = wavs[::-1]
    got = vf.restore_batch(pin_in, out=pin_out)
    assert got is pin_out and rel_rms(pin_out[::-1], out) < 1e-6      # replay of the cached graph, permuted batch
    with pytest.raises(ValueError):
        vf.restore_batch(wavs, mode=1)


@pytest.mark.timeout(900)
def test_restore_stream_matches_segmentwise_restore(home):
    """chunk = 30 s, no context: the stream's windows are the reference's segments and the output equals restore_inmem's.
    Short chunks with context: every sample comes out exactly once (length) and finite."""
    from voicefixer_b200 import synthetic, api
    vf = api.VoiceFixer(precision="tf32")
    wav = np.tile(synthetic.make_utterances(1, seconds=8.0, seed=81)[0], 5)[: 44100 * 33 + 77]      # 33 s: two segments
    whole = vf.restore_inmem(wav, cuda=True, mode=0)[0]
    blocks = [wav[i:i + 50000] for i in range(0, len(wav), 50000)]
    streamed = np.concatenate(list(vf.restore_stream(blocks, chunk_seconds=30.0, context_seconds=0.0)))
    assert streamed.shape == whole.shape and rel_rms(streamed, whole) < 1e-6
    short = np.concatenate(list(vf.restore_stream(blocks[:12], chunk_seconds=1.0, context_seconds=0.5)))
    assert short.shape[0] == sum(len(b) for b in blocks[:12]) and np.isfinite(short).all()


def test_too_small_workspace_is_refused_before_any_launch(states):
    """C ABI: a workspace smaller than the call needs returns VFX_ERR_WORKSPACE up front (no launch, nothing written)."""
    import ctypes
    import torch
    from voicefixer_b200 import _lib, synthetic
    from voicefixer_b200.engine import Engine
    eng = Engine(states[0], states[1], precision="tf32")
    wav = torch.from_numpy(synthetic.make_utterances(1, seconds=0.5, seed=3)).cuda()
    out = torch.full_like(wav, 7.0)
    need = eng.workspace_bytes(1, wav.shape[1])
    ws = torch.zeros(need // 4, dtype=torch.uint8, device="cuda")
    guard = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")          # whatever sits behind a too-small buffer
    launches = eng.launch_count()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = eng.lib.vfx_restore(eng.h, p(wav), 1, wav.shape[1], 0, None, p(out), p(ws), ws.numel(), None)
    torch.cuda.synchronize()
    assert rc == -3 and b"workspace too small" in eng.lib.vfx_last_error()
    assert eng.launch_count() == launches and float(out.min()) == 7.0 and int(guard.max()) == 0
    y = eng.restore(wav)                                                     # the engine is still usable
    assert bool(torch.isfinite(y).all())


@pytest.mark.timeout(600)
def test_long_recordings_are_chunked(home, monkeypatch):
    """restore_inmem batches the 30 s segments in chunks sized to the free memory; chunk size 1 (the reference's
    one-at-a-time walk) gives the same result as one batch."""
    from voicefixer_b200 import synthetic, api
    vf = api.VoiceFixer(precision="tf32")
    wav = np.tile(synthetic.make_utterances(1, seconds=6.0, seed=83)[0], 11)[: 44100 * 61]      # three segments
    assert vf._engine.max_batch(44100 * 30) >= 1
    whole = vf.restore_inmem(wav, cuda=True, mode=0)
    monkeypatch.setattr(vf._engine, "max_batch", lambda L, cap=64, reserve=0: 1)
    one_by_one = vf.restore_inmem(wav, cuda=True, mode=0)
    assert whole.shape == one_by_one.shape == (1, wav.shape[0]) and rel_rms(one_by_one, whole) < 1e-6
