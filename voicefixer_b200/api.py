"""Drop-in mirror of the reference's public API (voicefixer/base.py, voicefixer/vocoder/base.py).

Same class names, method signatures, checkpoint paths and error types; the bodies run the
hand-written sm_100a kernels through the C-ABI engine.  The `cuda` argument is accepted for
signature compatibility: the computation always runs on the GPU (there is no CPU path)."""
import os
import numpy as np
import torch

from .engine import Engine, default_precision
from . import wavio

SEG_LENGTH = 44100 * 30                      # voicefixer/base.py:116
ANALYSIS_CKPT = ".cache/voicefixer/analysis_module/checkpoints/vf.ckpt"
VOCODER_CKPT = ".cache/voicefixer/synthesis_module/44100/model.ckpt-1490000_trimed.pt"

_ERR0 = ("Error 0: The checkpoint for analysis module (vf.ckpt) is not found in "
         "~/.cache/voicefixer/analysis_module/checkpoints.")
_ERR1 = ("Error 1: The checkpoint for synthesis module / vocoder (model.ckpt-1490000_trimed) is not found in "
         "~/.cache/voicefixer/synthesis_module/44100.")


def _check_cuda(cuda):
    # tools/pytorch_util.py:6-8
    if not torch.cuda.is_available():
        if cuda:
            raise RuntimeError("Error: You set cuda=True but no cuda device found.")
        raise RuntimeError("voicefixer_b200 has no CPU path: a B200 (sm_100a) device is required.")


def _load_vocoder_state():
    path = os.path.join(os.path.expanduser("~"), VOCODER_CKPT)
    if not os.path.exists(path):
        raise RuntimeError(_ERR1)
    return torch.load(path, map_location="cpu")["generator"]


class _ModelShim:
    """Stands in for `VoiceFixer._model` (the reference nn.Module): callers poke
    `.parameters()`, `.to()`, `.eval()`, `.train()` and `.vocoder` (test/streamlit.py:40-42)."""

    def __init__(self, owner):
        self._owner = owner
        self.vocoder = owner._vocoder

    def parameters(self):
        return iter([self._owner._engine.arena])

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def train(self, mode=True):
        return self


class Vocoder:
    """voicefixer/vocoder/base.py:10-77."""

    def __init__(self, sample_rate, _engine=None, precision=None):
        if sample_rate != 44100:           # vocoder/config.py:28-31
            raise RuntimeError("Error: Vocoder currently only support 44100 samplerate.")
        self.rate = sample_rate
        if _engine is None:
            voc = _load_vocoder_state()
            _check_cuda(False)
            from .weights import pack_vocoder
            _engine = Engine(packed=pack_vocoder(voc, precision or default_precision()), precision=precision)
            self._standalone = True
        self._engine = _engine

    def forward(self, mel, cuda=False):
        """non-normalised mel [B, 1, T, 128] -> [B, 1, (T + T%2 + 4)*441] (vocoder/base.py:42-56)."""
        assert mel.size()[-1] == 128
        _check_cuda(cuda)
        out = self._engine.vocoder(mel[:, 0], input_is_log=False)
        out = out[:, None, :]
        return out if cuda else out.cpu()

    __call__ = forward

    def oracle(self, fpath, out_path, cuda=False):
        """vocoder/base.py:58-77: wav file -> |STFT| -> Slaney mel -> normalise -> Generator -> file."""
        _check_cuda(cuda)
        wav = wavio.read_wave(fpath, self.rate)[..., 0]
        cond = oracle_conditions(wav)
        out = self._engine.vocoder_cond(torch.from_numpy(cond), scale=2.0 ** 15)
        wavio.save_wave(out[:, None, :].cpu().numpy(), out_path, sample_rate=self.rate)


def oracle_conditions(wav):
    """Host front end of Vocoder.oracle (vocoder/base.py:61-73; numpy in the reference too):
    peak-normalise, |STFT| (2048/441/Hann/center, librosa-0.10 zero padding), Slaney-normalised
    HTK mel (librosa.filters.mel defaults), dB - 20, normalise, pad -> (1, Tc, 128) channels-last."""
    wav = np.asarray(wav, dtype=np.float32)
    wav = wav / np.max(np.abs(wav))
    n_fft, hop = 2048, 441
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)).astype(np.float32)
    x = np.pad(wav, (n_fft // 2, n_fft // 2), mode="constant")
    nfr = 1 + (len(x) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(nfr)[:, None]
    stft = np.abs(np.fft.rfft(x[idx] * win[None, :], axis=1).astype(np.complex64))      # (T, 1025)
    mel = stft @ wavio.slaney_htk_mel_basis().T                                         # (T, 128)
    min_level = np.exp(-100 / 20 * np.log(10))
    S = 20 * np.log10(np.maximum(min_level, np.abs(mel))) - 20
    S = np.clip(8.0 * ((S + 115.0) / 115.0) - 4.0, -4.0, 4.0).astype(np.float32)
    pad_tail = S.shape[0] % 2 + 4
    return np.concatenate([S, np.full((pad_tail, 128), -4.0, np.float32)], 0)[None]


class VoiceFixer:
    """voicefixer/base.py:10-146."""

    def __init__(self, precision=None):
        self.analysis_module_ckpt = os.path.join(os.path.expanduser("~"), ANALYSIS_CKPT)
        voc = _load_vocoder_state()        # the reference builds Vocoder first (restorer/model.py:180)
        if not os.path.exists(self.analysis_module_ckpt):
            raise RuntimeError(_ERR0)
        ana = torch.load(self.analysis_module_ckpt, map_location="cpu")
        _check_cuda(False)
        self._engine = Engine(ana, voc, precision=precision)
        self._vocoder = Vocoder(44100, _engine=self._engine)
        self._model = _ModelShim(self)

    @classmethod
    def from_engine(cls, engine):
        """The same API object around an engine whose weights are already on the device (e.g. an arena received through
        parallel.broadcast_arena on a non-zero rank: no checkpoint files are read)."""
        self = cls.__new__(cls)
        self.analysis_module_ckpt = os.path.join(os.path.expanduser("~"), ANALYSIS_CKPT)
        self._engine = engine
        self._vocoder = Vocoder(44100, _engine=engine)
        self._model = _ModelShim(self)
        return self

    # -- helpers kept from the reference API
    def _load_wav(self, path, sample_rate, threshold=0.95):
        return wavio.load_mono(path, sample_rate)                       # base.py:47-49

    def _trim_center(self, est, ref):                                   # base.py:63-76
        diff = abs(est.shape[-1] - ref.shape[-1])
        if est.shape[-1] == ref.shape[-1]:
            return est, ref
        if est.shape[-1] > ref.shape[-1]:
            min_len = min(est.shape[-1], ref.shape[-1])
            est = est[..., int(diff // 2): -int(diff // 2)]
            return est[..., :min_len], ref[..., :min_len]
        min_len = min(est.shape[-1], ref.shape[-1])
        ref = ref[..., int(diff // 2): -int(diff // 2)]
        return est[..., :min_len], ref[..., :min_len]

    def remove_higher_frequency(self, wav, ratio=0.95):                 # base.py:87-104
        out, _ = self._engine.hf_cut(torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32))[None], ratio)
        return out[0].cpu().numpy()

    @torch.no_grad()
    def restore_inmem(self, wav_10k, cuda=False, mode=0, your_vocoder_func=None, drop_masks_fn=None):
        """np (L,) float32 -> np (1, L) float32 (base.py:106-139).  The 30 s segments are
        independent, so all full segments run as one batch and the ragged tail as a
        second one; results are concatenated in order.

        drop_masks_fn (extension, mode 2 only): None = random Bernoulli(0.5) keep-masks like the reference's
        train-mode dropout; a callable (B, T) -> uint8 [2, B, T, 512] supplies them explicitly; False = no
        dropout (deterministic train-mode BN only)."""
        _check_cuda(cuda)
        if mode not in (0, 1, 2):
            raise ValueError("mode must be 0, 1 or 2")
        wav = np.ascontiguousarray(wav_10k, dtype=np.float32)
        n = wav.shape[0]
        segs = []
        bp = SEG_LENGTH
        while bp < n + SEG_LENGTH:                                       # base.py:117-119
            segs.append(wav[bp - SEG_LENGTH: bp])
            bp += SEG_LENGTH
        eng = self._engine
        dev = f"cuda:{eng.device}"
        groups = {}
        for i, s in enumerate(segs):
            groups.setdefault(len(s), []).append(i)
        outs = [None] * len(segs)
        for L, all_idxs in groups.items():
            Lnet = 512 * (L // 512) if mode == 1 else L                  # mode 1's pre-filter shortens the segment
            if mode == 2 and 1 + Lnet // 441 <= 64:
                # torch.nn.functional.batch_norm in the reference (train mode, 1x1 UNet centre)
                raise ValueError("Expected more than 1 value per channel when training, got input size "
                                 "torch.Size([1, 384, 1, 1])")
            # the reference walks the segments one at a time in constant memory; here they are batched, in chunks
            # sized to the device memory that is free right now (a 100-minute recording does not fit one launch sequence)
            chunk = eng.max_batch(L)
            for c0 in range(0, len(all_idxs), chunk):
                idxs = all_idxs[c0:c0 + chunk]
                x = torch.from_numpy(np.stack([segs[i] for i in idxs])).to(dev)
                if mode == 1:
                    x, _ = eng.hf_cut(x)                                 # shorter: 512*(L//512)
                masks = None
                if mode == 2 and drop_masks_fn is not False:
                    T = 1 + x.shape[1] // 441
                    # the reference's mode 2 is module.train(): both Dropout(0.5) of the denoiser are live and draw
                    # from torch's global RNG (base.py:114-115, restorer/model.py:76,90).  Same distribution here
                    # (keep-masks from torch's CPU generator, so torch.manual_seed controls it); the RNG *stream*
                    # differs from the reference's, which no reference test pins.  drop_masks_fn=False disables it.
                    masks = drop_masks_fn(x.shape[0], T) if drop_masks_fn else (torch.rand(2, x.shape[0], T, 512) >= 0.5)
                    masks = torch.as_tensor(masks).to(torch.uint8)
                    if tuple(masks.shape) != (2, x.shape[0], T, 512):
                        raise ValueError(f"dropout masks must have shape (2, {x.shape[0]}, {T}, 512), got {tuple(masks.shape)}")
                if your_vocoder_func is None:
                    y = eng.restore(x, mode=2 if mode == 2 else 0, drop_masks=masks)
                else:                                                    # base.py:126-129
                    mel = eng.frontend(x)
                    mel_log = eng.analysis(mel, mode=2 if mode == 2 else 0, drop_masks=masks)
                    denoised = (10 ** torch.clip(mel_log, max=5))[:, None]
                    ys = []
                    for k in range(denoised.shape[0]):                   # the hook sees what the reference hands it:
                        m = denoised[k:k + 1]                            # one segment [1, 1, T, 128], on the CPU unless cuda
                        yk = torch.as_tensor(your_vocoder_func(m if cuda else m.cpu())).to(dev).float()
                        if torch.max(torch.abs(yk)) > 1.0:               # base.py:131-133, per segment
                            yk = yk / torch.max(torch.abs(yk))
                            print("Warning: Exceed energy limit,", "input")
                        ys.append(self._trim_center(yk[:, 0], x[k:k + 1])[0])
                    y = torch.cat(ys, 0)
                for k, i in enumerate(idxs):
                    outs[i] = y[k]
        out = torch.cat(outs, -1)[None]
        return out.cpu().numpy()

    # ------------------------------------------------------------------ batch entry point (extension)
    def pinned_empty(self, shape):
        """numpy float32 array backed by pinned (page-locked) host memory: restore_batch copies such arrays to and from
        the GPU without an extra staging copy."""
        t = torch.empty(tuple(shape), dtype=torch.float32).pin_memory()
        a = t.numpy()
        self._pinned_keep = getattr(self, "_pinned_keep", [])
        self._pinned_keep.append(t)
        return a

    @torch.no_grad()
    def restore_batch(self, wavs, cuda=True, mode=0, out=None):
        """Extension of the reference API for many utterances of one length (the reference restores one array per call,
        voicefixer/base.py:106): wavs np (B, L) float32, L <= 30 s -> np (B, L) float32, every row exactly what
        restore_inmem(row, mode) returns (mode 0 here; modes 1 / 2 go through restore_inmem).  The launch sequence of a
        shape is captured into a CUDA graph on first use and replayed afterwards; host buffers move through pinned
        staging (directly, when `wavs` / `out` are pinned: see pinned_empty).  Without `out`, the result is a view of an
        internal pinned buffer that the next call with the same shape overwrites."""
        _check_cuda(cuda)
        if mode != 0:
            raise ValueError("restore_batch runs mode 0; use restore_inmem for modes 1 and 2")
        x = np.ascontiguousarray(wavs, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] > SEG_LENGTH:
            raise ValueError("restore_batch takes (B, L) with L <= 30 s (longer inputs: restore_inmem segments them)")
        eng = self._engine
        dev = f"cuda:{eng.device}"
        plans = self.__dict__.setdefault("_batch_plans", {})
        key = tuple(x.shape)
        if key not in plans:
            if len(plans) >= 4:                                            # a few shapes at most: graphs own workspaces
                plans.pop(next(iter(plans)))
            p = {"dev_in": torch.empty(key, device=dev), "dev_out": torch.empty(key, device=dev),
                 "host_in": torch.empty(key).pin_memory(), "host_out": torch.empty(key).pin_memory()}
            p["dev_in"].copy_(torch.from_numpy(x))
            p["graph"] = eng.make_graph(p["dev_in"], p["dev_out"], mode=0)
            plans[key] = p
        p = plans[key]
        src = torch.from_numpy(x)
        if not src.is_pinned():
            p["host_in"].copy_(src)
            src = p["host_in"]
        p["dev_in"].copy_(src, non_blocking=True)
        p["graph"].replay()
        dst = p["host_out"]
        if out is not None:
            o = torch.from_numpy(out)
            if o.is_pinned() and o.is_contiguous() and tuple(o.shape) == key and o.dtype == torch.float32:
                dst = o
        dst.copy_(p["dev_out"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if out is not None and dst is p["host_out"]:
            out[...] = dst.numpy()
            return out
        return out if out is not None else dst.numpy()

    # ------------------------------------------------------------------ streaming entry point (extension)
    @torch.no_grad()
    def restore_stream(self, blocks, chunk_seconds=2.0, context_seconds=1.0, cuda=True):
        """Low-latency form of restore_inmem (mode 0) for audio that arrives in blocks: a generator that takes an iterable
        of 1-D float32 blocks of any sizes (44.1 kHz mono) and yields restored blocks as soon as they can be computed.

        The reference has no streaming path: it cuts a recording into hard 30 s segments whose BiGRU and UNet see the whole
        segment (voicefixer/base.py:116-138).  Here every emitted chunk of `chunk_seconds` is restored inside a window that
        also holds `context_seconds` of already-received audio on the left and of look-ahead on the right
        (window = [pos - ctx, pos + chunk + ctx), clipped to the stream), and only the chunk itself is kept.  The
        algorithmic latency is chunk + context seconds of audio plus one B = 1 launch sequence (see bench.py
        `workloads.stream`).  With chunk_seconds = 30 and context_seconds = 0 the windows ARE the reference's segments
        and the concatenated output equals restore_inmem's exactly; smaller windows trade fidelity to whole-segment
        processing for latency (how much depends on the trained weights' context dependence; it cannot be pinned with
        the synthetic checkpoints available offline).  The total output length equals the input length."""
        _check_cuda(cuda)
        chunk = int(round(chunk_seconds * 44100))
        ctx = int(round(context_seconds * 44100))
        if chunk <= 0 or ctx < 0:
            raise ValueError("chunk_seconds must be positive and context_seconds non-negative")
        eng = self._engine
        dev = f"cuda:{eng.device}"
        buf = np.zeros(0, dtype=np.float32)       # samples [base, base + len(buf)) of the stream
        base = 0                                  # stream index of buf[0]
        pos = 0                                   # next sample to emit

        def run(lo, hi, keep_lo, keep_hi):
            x = torch.from_numpy(np.ascontiguousarray(buf[lo - base: hi - base]))[None].to(dev)
            y = eng.restore(x, mode=0)
            return y[0, keep_lo - lo: keep_hi - lo].cpu().numpy()

        for block in blocks:
            block = np.asarray(block, dtype=np.float32).reshape(-1)
            buf = np.concatenate([buf, block])
            while base + len(buf) >= pos + chunk + ctx:                  # a full chunk plus its look-ahead has arrived
                lo = max(0, pos - ctx)
                yield run(lo, pos + chunk + ctx, pos, pos + chunk)
                pos += chunk
                drop = max(0, pos - ctx) - base                          # audio older than the next window's left context
                if drop > 0:
                    buf, base = buf[drop:], base + drop
        end = base + len(buf)
        while pos < end:                                                 # flush: the last windows have no (full) look-ahead
            hi = min(end, pos + chunk + ctx)
            keep_hi = min(end, pos + chunk)
            if end - keep_hi <= 1024 and end - keep_hi > 0 and ctx == 0:  # never leave a tail the front end cannot pad
                keep_hi = hi = end
            lo = max(0, pos - ctx)
            if hi - lo <= 1024:                                          # reflect padding needs > 1024 samples (base.py:78)
                lo = max(0, hi - 1025 - ctx)
                if hi - lo <= 1024:
                    raise RuntimeError("restore_stream: fewer than 1025 samples in total; the front end reflect-pads 1024")
            yield run(lo, hi, pos, keep_hi)
            pos = keep_hi

    def restore(self, input, output, cuda=False, mode=0, your_vocoder_func=None):
        wav_10k = self._load_wav(input, sample_rate=44100)               # base.py:141-146
        out_np_wav = self.restore_inmem(wav_10k, cuda=cuda, mode=mode, your_vocoder_func=your_vocoder_func)
        wavio.save_wave(out_np_wav, fname=output, sample_rate=44100)
