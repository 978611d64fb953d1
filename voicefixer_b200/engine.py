"""Python wrapper of the C-ABI engine: device memory, streams and weight upload (plumbing only).

All compute goes through libvfx_b200.so; torch is used for device buffers, the current CUDA
stream and (in parallel.py) torch.distributed."""
import ctypes
import os
import numpy as np
import torch

from . import _lib
from .weights import pack_analysis, pack_vocoder

ALIGN = 256


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def default_precision():
    """"tf32": tcgen05 kind::tf32 on tf32-rounded fp32 operands -- the arithmetic class of the reference's own CUDA path
    (cuDNN TF32 convolutions; waveform within 2e-3 relative RMS of the fp32 CPU path).  VFX_PRECISION selects another mode:
    "bf16" (tcgen05 kind::f16, 1.6x faster, 1.3e-2 relative RMS / 2.7e-3 mean-abs: inside the reference's own 1e-2 mean-abs
    acceptance bar), "fp16" (tcgen05 kind::f16 on fp16 operands: tf32's 10-bit mantissa, hence tf32's parity, at the bf16 mode's
    speed -- within fp16's exponent range) or "fp32" (SIMT fp32 validation path, reference-exact to ~4e-6)."""
    return os.environ.get("VFX_PRECISION", "tf32")


class Planner:
    """Planning-only engine (vfx_engine_create with device -1): validates a packed weight set against what the
    kernels expect (names, byte sizes) and sizes the activation workspace for a batch -- no GPU, no CUDA call.
    `packed`: {name: CPU tensor} from weights.pack_analysis / pack_vocoder (either both, or the vocoder alone)."""

    def __init__(self, packed, precision):
        self.lib = _lib.load()
        self.precision = precision
        h = ctypes.c_void_p()
        _lib.check(self.lib.vfx_engine_create(ctypes.byref(h), -1, _lib.PREC[precision]), "vfx_engine_create")
        self.h = h
        table, self.weight_bytes = Engine.layout(packed)
        for name, off, nbytes in table:          # addresses are never dereferenced by a planning-only engine
            _lib.check(self.lib.vfx_engine_set_tensor(self.h, name.encode(), ctypes.c_void_p(ALIGN + off), nbytes),
                       f"set_tensor({name})")
        _lib.check(self.lib.vfx_engine_finalize(self.h), "vfx_engine_finalize")

    def workspace_bytes(self, B, L):
        """restore() on B items of L samples; 0 for a vocoder-only weight set."""
        return int(self.lib.vfx_workspace_bytes(self.h, int(B), int(L)))

    def workspace_bytes_frames(self, B, T):
        """analysis / vocoder entry points on B items of T mel frames."""
        return int(self.lib.vfx_workspace_bytes_frames(self.h, int(B), int(T)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vfx_engine_destroy(self.h)
            self.h = None


class Engine:
    """One engine per process / GPU.  `ana`, `voc`: reference-layout state dicts (CPU tensors)."""

    def __init__(self, ana=None, voc=None, device=None, precision=None, packed=None):
        if not torch.cuda.is_available():
            raise RuntimeError("voicefixer_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.precision = precision or default_precision()
        if self.precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {list(_lib.PREC)}")
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vfx_engine_create(ctypes.byref(h), self.device, _lib.PREC[self.precision]),
                       "vfx_engine_create")
        self.h = h
        self._ws = None
        self.arena = None
        self.table = None
        if packed is None and ana is not None:
            packed = {}
            packed.update(pack_analysis(ana, self.precision))
            packed.update(pack_vocoder(voc, self.precision))
        if packed is not None:
            self.upload(packed)

    # ------------------------------------------------------------------ weights
    @staticmethod
    def layout(packed):
        """[(name, offset, nbytes)] and total bytes of the single weight arena."""
        table, off = [], 0
        for name in sorted(packed):
            t = packed[name]
            nbytes = t.numel() * t.element_size()
            table.append((name, off, nbytes))
            off += (nbytes + ALIGN - 1) // ALIGN * ALIGN
        return table, off

    def upload(self, packed):
        table, total = self.layout(packed)
        host = torch.empty(total, dtype=torch.uint8).pin_memory()
        for name, off, nbytes in table:
            host[off:off + nbytes] = packed[name].contiguous().view(torch.uint8).reshape(-1)
        arena = torch.empty(total, dtype=torch.uint8, device=f"cuda:{self.device}")
        arena.copy_(host, non_blocking=False)
        self.attach(arena, table)

    def attach(self, arena, table):
        """Registers views of an arena that is already on this device (e.g. after an NCCL broadcast)."""
        self.arena, self.table = arena, table
        base = arena.data_ptr()
        for name, off, nbytes in table:
            _lib.check(self.lib.vfx_engine_set_tensor(self.h, name.encode(), ctypes.c_void_p(base + off), nbytes),
                       f"set_tensor({name})")
        _lib.check(self.lib.vfx_engine_finalize(self.h), "vfx_engine_finalize")

    def set_option(self, key, value):
        _lib.check(self.lib.vfx_engine_set_option(self.h, key.encode(), int(value)), "set_option")

    def profile(self, on=True):
        self.set_option("profile", 1 if on else 0)

    def profile_report(self):
        """{tag: dict(count, ms, flops, bytes)} of everything launched since the last report."""
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(self.lib.vfx_profile_report(self.h, buf, len(buf)), "vfx_profile_report")
        rep = {}
        for line in buf.value.decode().splitlines():
            tag, n, ms, fl, by = line.split()
            rep[tag] = dict(count=int(n), ms=float(ms), flops=float(fl), bytes=float(by))
        return rep

    def launch_count(self):
        return int(self.lib.vfx_launch_count())

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.vfx_engine_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ workspace
    def _workspace(self, nbytes):
        if nbytes == 0:
            raise _lib.VfxError("workspace query returned 0: the shape is out of range (restore() needs L > 1024 samples) or this "
                                "engine holds the vocoder weights only (stand-alone Vocoder: restore/analysis are unavailable)")
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None     # (captured graphs own their workspace tensors; see make_graph)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{self.device}")
        return self._ws

    def max_batch(self, L, cap=64, reserve=2 << 30):
        """Largest number of equal-length items (<= cap) whose restore() workspace fits the memory that is free right now
        (plus the workspace this engine already holds), keeping `reserve` bytes back.  At least 1."""
        free, _ = torch.cuda.mem_get_info(self.device)
        avail = free + (self._ws.numel() if self._ws is not None else 0) - reserve
        per_item = max(1, self.workspace_bytes(2, L) - self.workspace_bytes(1, L))
        return int(max(1, min(cap, avail // per_item)))

    def workspace_bytes(self, B, L):
        return int(self.lib.vfx_workspace_bytes(self.h, B, L))

    def _dev(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        return x.to(device=f"cuda:{self.device}", dtype=torch.float32).contiguous()

    # ------------------------------------------------------------------ the five seams
    def frontend(self, wav, return_sp=False):
        """wav (B, L) -> mel (B, T, 128) [, sp (B, T, 1025)]."""
        wav = self._dev(wav)
        B, L = wav.shape
        T = 1 + L // 441
        mel = torch.empty(B, T, 128, device=wav.device)
        sp = torch.empty(B, T, 1025, device=wav.device) if return_sp else None
        _lib.check(self.lib.vfx_frontend_mel(self.h, _ptr(wav), B, L, _ptr(mel), _ptr(sp), _stream()), "vfx_frontend_mel")
        return (mel, sp) if return_sp else mel

    def analysis(self, mel, mode=0, drop_masks=None):
        """mel (B, T, 128) linear -> log10 mel (B, T, 128)."""
        mel = self._dev(mel)
        B, T, F = mel.shape
        if F != 128:
            raise AssertionError("mel must have 128 bins")
        out = torch.empty_like(mel)
        ws = self._workspace(int(self.lib.vfx_workspace_bytes_frames(self.h, B, T)))
        dm = None
        if drop_masks is not None:
            dm = drop_masks.to(device=mel.device, dtype=torch.uint8).contiguous()
            assert tuple(dm.shape) == (2, B, T, 512)
        _lib.check(self.lib.vfx_analysis(self.h, _ptr(mel), B, T, int(mode), _ptr(dm), _ptr(out), _ptr(ws), ws.numel(),
                                         _stream()), "vfx_analysis")
        return out

    def vocoder(self, mel, input_is_log=False, trim_len=-1, scale=1.0):
        """mel (B, T, 128) -> wav (B, out_len)."""
        mel = self._dev(mel)
        B, T, F = mel.shape
        if F != 128:
            raise AssertionError("mel must have 128 bins")
        S = (T + T % 2 + 4) * 441
        out = torch.empty(B, S if trim_len < 0 else trim_len, device=mel.device)
        ws = self._workspace(int(self.lib.vfx_workspace_bytes_frames(self.h, B, T)))
        _lib.check(self.lib.vfx_vocoder(self.h, _ptr(mel), B, T, int(bool(input_is_log)), _ptr(out), int(trim_len),
                                        float(scale), _ptr(ws), ws.numel(), _stream()), "vfx_vocoder")
        return out

    def vocoder_cond(self, cond, trim_len=-1, scale=1.0):
        """normalised conditions (B, Tc, 128) [channels-last] -> wav (B, Tc*441)."""
        cond = self._dev(cond)
        B, Tc, F = cond.shape
        out = torch.empty(B, Tc * 441 if trim_len < 0 else trim_len, device=cond.device)
        ws = self._workspace(int(self.lib.vfx_workspace_bytes_frames(self.h, B, Tc)))
        _lib.check(self.lib.vfx_vocoder_cond(self.h, _ptr(cond), B, Tc, _ptr(out), int(trim_len), float(scale), _ptr(ws),
                                             ws.numel(), _stream()), "vfx_vocoder_cond")
        return out

    def restore(self, wav, mode=0, drop_masks=None, out=None, ws=None):
        """wav (B, L) on device -> restored wav (B, L) on device (one launch sequence)."""
        wav = self._dev(wav)
        B, L = wav.shape
        if out is None:
            out = torch.empty_like(wav)
        if ws is None:
            ws = self._workspace(self.workspace_bytes(B, L))
        dm = None
        if drop_masks is not None:
            dm = drop_masks.to(device=wav.device, dtype=torch.uint8).contiguous()
        _lib.check(self.lib.vfx_restore(self.h, _ptr(wav), B, L, int(mode), _ptr(dm), _ptr(out), _ptr(ws), ws.numel(),
                                        _stream()), "vfx_restore")
        return out

    def make_graph(self, wav, out, mode=0):
        """Captures the whole restore() launch sequence (~360 kernels) for the given device buffers into a
        CUDA graph; returns an object whose .replay() re-runs it (inputs are read from `wav`, results land in
        `out`).  Shapes, buffers and the workspace are frozen into the graph."""
        assert wav.is_cuda and out.is_cuda and wav.shape == out.shape
        # the graph bakes raw device pointers in: it gets its OWN workspace, and the returned object keeps every
        # buffer it touches alive (a later, larger call may replace the engine's shared workspace)
        ws = torch.empty(self.workspace_bytes(*wav.shape), dtype=torch.uint8, device=wav.device)
        side = torch.cuda.Stream(device=wav.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                                # warm-up: one-time attribute / table setup
            for _ in range(2):
                self.restore(wav, mode=mode, out=out, ws=ws)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.restore(wav, mode=mode, out=out, ws=ws)
        g._keep = (ws, wav, out, self.arena)
        return g

    def hf_cut(self, wav, ratio=0.95):
        wav = self._dev(wav)
        B, L = wav.shape
        out = torch.empty(B, 512 * (L // 512), device=wav.device)
        cut = torch.empty(B, dtype=torch.int32, device=wav.device)
        nfr = 1 + L // 512
        ws = self._workspace(B * nfr * (1025 * 8 + 2048 * 4) + B * 1025 * 8 + (1 << 16))
        _lib.check(self.lib.vfx_hf_cut(self.h, _ptr(wav), B, L, float(ratio), _ptr(out), _ptr(cut), _ptr(ws), ws.numel(),
                                       _stream()), "vfx_hf_cut")
        return out, cut
