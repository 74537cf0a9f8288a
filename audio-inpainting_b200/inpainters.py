"""Drop-in counterparts of the reference's three NMF inpainting classes, running on the B200 path.

Same class names, method names, argument meaning, return values and attribute names as
    NMFFairGapInpainter   (main4_NMF_gap.py:11-84)
    NMFFairInpainter      (main4_NMF_mask.py:11-89)
    SpectralInpainter     (main4_NMF.py:27-161)
so the module tails of those scripts (main4_NMF_gap.py:86-89, main4_NMF_mask.py:92-95, main4_NMF.py:163-172)
run unchanged against them.  The constants the reference hard-codes are keyword arguments of the constructors.
Differences: PNG rendering (matplotlib) is not part of the hot path and is skipped; printing is optional.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from scipy.io import wavfile

from . import ops


def _device(device):
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError("ainmf runs on CUDA devices only (no CPU fallback)")
    return d


class _ColumnInpainter:
    """Shared body of NMFFairGapInpainter / NMFFairInpainter (the two scripts differ in two constants)."""
    _threshold = 1e-4
    _frac = (9, 10)
    _out_name = "fixed_nmf_gap.wav"
    _out_dir = "demo_assets/part2"

    def __init__(self, filename, device="cuda", n_fft=1024, hop_length=256, seed=42, max_iter=200, tol=1e-4,
                 threshold=None, frac=None, output_dir=None, verbose=False):
        self.filename = filename
        self.sr = None
        self.signal = None
        self.device = _device(device)
        self.n_fft, self.hop_length = n_fft, hop_length
        self.seed, self.max_iter, self.tol = seed, max_iter, tol
        self.threshold = self._threshold if threshold is None else threshold
        self.frac = self._frac if frac is None else frac
        self.output_dir = self._out_dir if output_dir is None else output_dir
        self.verbose = verbose
        self._x = None          # device copy of self.signal, [1, N]
        # results of the last restore() for inspection (the reference keeps them as locals)
        self.bad_cols_ = None
        self.n_iter_ = None
        self.reconstruction_err_ = None
        self.W_ = None
        self.H_ = None

    # main4_NMF_gap.py:17-26
    def load_damaged_data(self):
        if not os.path.exists(self.filename):
            print("File not found! Please run the data generator first")
            return
        self.sr, data = wavfile.read(self.filename)
        if data.dtype == np.int16:
            pcm = torch.from_numpy(np.ascontiguousarray(data)).to(self.device)
            pcm = pcm.unsqueeze(0)
            x, _ = ops.load_pcm16(pcm)
        else:
            if data.ndim > 1:
                data = data.mean(axis=1)
            x = torch.from_numpy(data.astype(np.float32)).to(self.device).unsqueeze(0)
            peak = x.abs().max()
            if float(peak) > 0:
                x = x / peak
        self._x = x.contiguous()
        self.signal = self._x[0].cpu().numpy()
        if self.verbose:
            print(f"NMF loaded damaged audio: {len(self.signal)} samples")

    def _ensure_device_signal(self):
        if self._x is None or self._x.shape[1] != len(self.signal):
            self._x = torch.from_numpy(np.ascontiguousarray(self.signal, np.float32)).to(self.device).unsqueeze(0)
        return self._x

    # main4_NMF_gap.py:28-40 / main4_NMF_mask.py:28-45
    def _mask(self, n_frames, hop_length):
        x = self._ensure_device_signal()
        _, idx, nb = ops.gap_mask(x, hop_length, n_frames, self.threshold, self.frac[0], self.frac[1])
        n = int(nb[0])
        return idx[0, :n].cpu().numpy().astype(np.int64)

    # main4_NMF_gap.py:42-72
    def restore(self, n_components=40, n_iter=50):
        """`n_iter` is accepted and ignored, exactly as in the reference (its budget is max_iter=200)."""
        if self.signal is None:
            return
        x = self._ensure_device_signal()
        y, idx, nb, W, H, err, nit = ops.nmf_inpaint(
            x, self.n_fft, self.hop_length, n_components, self.max_iter, self.tol, self.seed, self.threshold,
            self.frac[0], self.frac[1], -1, -1, 1, None, None)
        n = int(nb[0])
        self.bad_cols_ = idx[0, :n].cpu().numpy().astype(np.int64)
        if self.verbose:
            print(f"Detected {n} damaged spectrum columns (~{n * self.hop_length / self.sr:.2f}s)")
        if n == 0:
            return self.signal
        self.n_iter_ = int(nit[0])
        self.reconstruction_err_ = float(err[0])
        self.W_, self.H_ = W[0], H[0]
        return y[0].cpu().numpy()

    # main4_NMF_gap.py:74-77 (PNG part omitted)
    def save_result(self, audio):
        os.makedirs(self.output_dir, exist_ok=True)
        path = os.path.join(self.output_dir, self._out_name)
        y = torch.from_numpy(np.ascontiguousarray(audio, np.float32)).to(self.device)
        wavfile.write(path, self.sr, ops.store_pcm16(y).cpu().numpy())
        if self.verbose:
            print(f"NMF restoration complete ({path})")


class NMFFairGapInpainter(_ColumnInpainter):
    """main4_NMF_gap.py:11.  threshold 1e-4, fraction 0.9."""

    def get_gap_mask(self, n_frames, hop_length):
        return self._mask(n_frames, hop_length)


class NMFFairInpainter(_ColumnInpainter):
    """main4_NMF_mask.py:11.  threshold 0.01, fraction 0.8."""
    _threshold = 0.01
    _frac = (4, 5)
    _out_name = "fixed_nmf_random.wav"
    _out_dir = "demo_assets"

    def get_mask_from_signal(self, n_frames, hop_length):
        return self._mask(n_frames, hop_length)


class SpectralInpainter:
    """main4_NMF.py:27.  Short segment, known gap, 50 refits."""

    def __init__(self, filename, duration=0.1, device="cuda", n_fft=512, hop_length=128, seed=0, max_iter=200,
                 tol=1e-4, output_dir="demo_assets/part0", verbose=False):
        self.filename = filename
        self.duration = duration
        self.sr = None
        self.raw_audio = None
        self.restored_audio = None
        self.corrupted_audio = None
        self.device = _device(device)
        self.n_fft, self.hop_length = n_fft, hop_length
        self.seed, self.max_iter, self.tol = seed, max_iter, tol
        self.output_dir = output_dir
        self.verbose = verbose
        self.n_iter_ = None
        self.reconstruction_err_ = None
        self.cols_ = None
        self.snr_ = None
        self.local_snr_ = None

    # main4_NMF.py:35-45 (host: one pass over the file, outside the op)
    def load_data(self):
        self.sr, data = wavfile.read(self.filename)
        if data.dtype != np.float32:
            data = data.astype(np.float32) / np.iinfo(data.dtype).max
        if len(data.shape) > 1:
            data = data.mean(axis=1)
        data = data / np.max(np.abs(data)) if np.max(np.abs(data)) > 0 else data
        n = int(self.duration * self.sr)
        start = len(data) // 2
        self.raw_audio = data[start:start + n]
        if self.verbose:
            print(f"Audio loaded: {len(self.raw_audio)} samples")

    # main4_NMF.py:47-60
    def apply_mask(self, gap_ratio=0.2):
        n = len(self.raw_audio)
        self.gap_start = int(n * 0.4)
        self.gap_end = int(self.gap_start + n * gap_ratio)
        self.corrupted_audio = self.raw_audio.copy()
        fade_len = min(100, self.gap_start, n - self.gap_end)
        if fade_len > 0:
            window = np.linspace(1, 0, fade_len)
            self.corrupted_audio[self.gap_start - fade_len:self.gap_start] *= window
            self.corrupted_audio[self.gap_end:self.gap_end + fade_len] *= window[::-1]
        self.corrupted_audio[self.gap_start:self.gap_end] = 0
        return self.gap_start, self.gap_end

    def _gap_columns(self):
        """col_start/col_end of main4_NMF.py:74-76 with scipy's frame times ($SP/scipy/signal/_spectral_py.py:2324-2327)."""
        n, hop, fs = self.n_fft, self.hop_length, float(self.sr)
        t = np.array([n / 2, n / 2 + hop]) / fs
        t -= (n / 2) / fs
        t_step = t[1] - t[0]
        return int(self.gap_start / self.sr / t_step), int(self.gap_end / self.sr / t_step)

    # main4_NMF.py:62-112
    def restore_with_nmf(self, n_components=30, n_iter=20):
        cs, ce = self._gap_columns()
        self.cols_ = (cs, ce)
        x = torch.from_numpy(np.ascontiguousarray(self.corrupted_audio, np.float32)).to(self.device).unsqueeze(0)
        y, _, _, _, _, err, nit = ops.nmf_inpaint(x, self.n_fft, self.hop_length, n_components, self.max_iter,
                                                  self.tol, self.seed, 1e-4, 9, 10, cs, ce, n_iter, None, None)
        self.n_iter_ = int(nit[0])
        self.reconstruction_err_ = float(err[0])
        raw = torch.from_numpy(np.ascontiguousarray(self.raw_audio, np.float32)).to(self.device)
        final = self._blend_boundaries(raw, y[0][:len(self.raw_audio)].contiguous())
        self.snr_ = ops.snr_db(raw, final)                                          # main4_NMF.py:99-102
        self.local_snr_ = ops.snr_db(raw, final, self.gap_start, self.gap_end)      # :104-108
        self.restored_audio = final.cpu().numpy()
        if self.verbose:
            print(f"SNR: {self.snr_:.2f} dB, Local SNR: {self.local_snr_:.2f} dB")
        return self.restored_audio

    # main4_NMF.py:114-126 (on the device: ops.blend_boundaries reproduces numpy's float64 ramp arithmetic exactly)
    def _blend_boundaries(self, raw, restored):
        return ops.blend_boundaries(raw, restored, int(self.gap_start), int(self.gap_end), 50)

    # main4_NMF.py:128-137 (WAV part)
    def save_results(self):
        os.makedirs(self.output_dir, exist_ok=True)
        for name, audio in (("nmf_corrupted.wav", self.corrupted_audio), ("nmf_restored.wav", self.restored_audio),
                            ("nmf_original.wav", self.raw_audio)):
            y = torch.from_numpy(np.ascontiguousarray(audio, np.float32)).to(self.device)
            wavfile.write(os.path.join(self.output_dir, name), self.sr, ops.store_pcm16(y).cpu().numpy())

    def visualize(self):
        """Plotting is outside the hot path (main4_NMF.py:139-161); intentionally a no-op."""
        return None
