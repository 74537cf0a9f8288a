"""The reference's glue restated around the SAME scipy / sklearn calls -- TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows.  The constants the reference hard-codes
(n_fft, hop, threshold, fraction, seed, max_iter) are arguments here so that the same code
serves configs 1-5 of BASELINE.json.  This module is what ``bench.py`` times as the CPU baseline.
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy import signal

try:  # sklearn warns at max_iter; the reference lets the warning print
    from sklearn.exceptions import ConvergenceWarning
except Exception:  # pragma: no cover
    ConvergenceWarning = Warning


def load_normalised(data: np.ndarray) -> np.ndarray:
    """main4_NMF_gap.py:21-24 / main4_NMF_mask.py:21-24: mono mean, float32, x / max|x|."""
    if data.ndim > 1:
        data = data.mean(axis=1)
    data = data.astype(np.float32)
    peak = np.max(np.abs(data))
    if peak > 0:
        data = data / peak
    return data


def quantise_int16(audio: np.ndarray) -> np.ndarray:
    """save_result / save_wav: main4_NMF_gap.py:76-77, main4_NMF.py:23-24 (truncation toward 0)."""
    audio = np.clip(audio, -1.0, 1.0)
    return (audio * 32767).astype(np.int16)


def column_mask(x: np.ndarray, n_frames: int, hop: int, threshold: float, frac: float) -> np.ndarray:
    """get_gap_mask (main4_NMF_gap.py:28-40, thr 1e-4, frac 0.9) and get_mask_from_signal
    (main4_NMF_mask.py:28-45, thr 0.01, frac 0.8)."""
    is_gap = np.abs(x) < threshold
    bad = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # empty slice at the last column -> nan -> not bad
        for col in range(n_frames):
            c = col * hop
            ws = max(0, c - hop // 2)
            we = min(len(x), c + hop // 2)
            if np.mean(is_gap[ws:we]) > frac:
                bad.append(col)
    return np.array(bad, dtype=np.int64)


def stft_mag_phase(x, sr, n_fft, hop):
    """main4_NMF_gap.py:47-49."""
    _, t, Z = signal.stft(x, sr, nperseg=n_fft, noverlap=n_fft - hop)
    return Z, np.abs(Z), np.angle(Z), t


def impute(mag, bad_cols):
    """main4_NMF_gap.py:55-59 (the O(T^2) list build is replaced by a boolean mask: same set)."""
    good = np.ones(mag.shape[1], dtype=bool)
    good[bad_cols] = False
    cur = mag.copy()
    avg = np.mean(mag[:, good], axis=1, keepdims=True)
    cur[:, bad_cols] = avg
    return cur


def nmf_fit(X, K, seed=None, max_iter=200, tol=1e-4, W0=None, H0=None, solver="cd",
            beta_loss="frobenius"):
    """NMF(n_components=K, init='random', random_state=seed, max_iter=200).fit_transform
    (main4_NMF_gap.py:62-64).  With W0/H0 given, init='custom' on copies (sklearn updates
    contiguous inputs in place)."""
    from sklearn.decomposition import NMF
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", ConvergenceWarning)
        if W0 is not None:
            m = NMF(n_components=K, init="custom", max_iter=max_iter, tol=tol, solver=solver,
                    beta_loss=beta_loss)
            W = m.fit_transform(X, W=np.array(W0, dtype=X.dtype, order="C", copy=True),
                                H=np.array(H0, dtype=X.dtype, order="C", copy=True))
        else:
            m = NMF(n_components=K, init="random", random_state=seed, max_iter=max_iter, tol=tol,
                    solver=solver, beta_loss=beta_loss)
            W = m.fit_transform(X)
    return W, m.components_, int(m.n_iter_), float(m.reconstruction_err_)


def restore_columns(x, sr, *, n_fft=1024, hop=256, threshold=1e-4, frac=0.9, K=40, seed=42,
                    max_iter=200, tol=1e-4, W0=None, H0=None, solver="cd", beta_loss="frobenius",
                    return_all=False):
    """NMFFairGapInpainter.restore (main4_NMF_gap.py:42-72) / NMFFairInpainter.restore
    (main4_NMF_mask.py:47-77) with the constants as arguments."""
    Z, mag, phase, _ = stft_mag_phase(x, sr, n_fft, hop)
    bad = column_mask(x, mag.shape[1], hop, threshold, frac)
    if len(bad) == 0:
        return (x, dict(bad=bad)) if return_all else x
    cur = impute(mag, bad)
    W, H, n_iter, err = nmf_fit(cur, K, seed, max_iter, tol, W0, H0, solver, beta_loss)
    V_hat = W @ H
    final = mag.copy()
    final[:, bad] = V_hat[:, bad]
    Zr = final * np.exp(1j * phase)
    _, y = signal.istft(Zr, sr, nperseg=n_fft, noverlap=n_fft - hop)
    y = y[: len(x)]
    if return_all:
        return y, dict(bad=bad, Z=Z, mag=mag, X=cur, W=W, H=H, n_iter=n_iter, err=err, Zr=Zr)
    return y


# ---- Part 0 (main4_NMF.py) ----------------------------------------------------------------------

def part0_load(data: np.ndarray, sr: int, duration: float) -> np.ndarray:
    """SpectralInpainter.load_data (main4_NMF.py:35-45) on already-read wav samples."""
    if data.dtype != np.float32:
        data = data.astype(np.float32) / np.iinfo(data.dtype).max
    if data.ndim > 1:
        data = data.mean(axis=1)
    peak = np.max(np.abs(data))
    data = data / peak if peak > 0 else data
    n = int(duration * sr)
    start = len(data) // 2
    return data[start:start + n]


def part0_apply_mask(raw: np.ndarray, gap_ratio=0.2):
    """SpectralInpainter.apply_mask (main4_NMF.py:47-60)."""
    n = len(raw)
    gs = int(n * 0.4)
    ge = int(gs + n * gap_ratio)
    cor = raw.copy()
    fade = min(100, gs, n - ge)
    if fade > 0:
        w = np.linspace(1, 0, fade)
        cor[gs - fade:gs] *= w
        cor[ge:ge + fade] *= w[::-1]
    cor[gs:ge] = 0
    return cor, gs, ge


def part0_blend(raw, restored, gs, ge, blend_width=50):
    """SpectralInpainter._blend_boundaries (main4_NMF.py:114-126)."""
    final = raw.copy()
    m = np.linspace(0, 1, blend_width)
    final[gs:ge] = restored[gs:ge]
    final[gs - blend_width:gs] = final[gs - blend_width:gs] * (1 - m) + restored[gs - blend_width:gs] * m
    final[ge:ge + blend_width] = final[ge:ge + blend_width] * m + restored[ge:ge + blend_width] * (1 - m)
    return final


def part0_restore(raw, corrupted, sr, gs, ge, *, n_fft=512, hop=128, K=40, n_outer=50, seed=0,
                  max_iter=200, tol=1e-4, return_all=False):
    """SpectralInpainter.restore_with_nmf (main4_NMF.py:62-112) incl. blend and SNR."""
    _, t, Z = signal.stft(corrupted, sr, nperseg=n_fft, noverlap=n_fft - hop)
    mag, phase = np.abs(Z), np.angle(Z)
    t_step = t[1] - t[0]
    cs = int(gs / sr / t_step)
    ce = int(ge / sr / t_step)
    cur = mag.copy()
    cur[:, cs:ce] = np.mean(mag[:, :cs], axis=1, keepdims=True)
    n_iters, err = [], 0.0
    for _ in range(n_outer):
        W, H, it, err = nmf_fit(cur, K, seed, max_iter, tol)
        n_iters.append(it)
        cur[:, cs:ce] = (W @ H)[:, cs:ce]
    Zr = cur * np.exp(1j * phase)
    _, y = signal.istft(Zr, sr, nperseg=n_fft, noverlap=n_fft - hop)
    y = y[: len(raw)]
    pre_blend = y
    y = part0_blend(raw, y, gs, ge)
    if return_all:
        return y, dict(cols=(cs, ce), n_iters=n_iters, err=err, pre_blend=pre_blend, mag=cur)
    return y


def snr_db(ref, est):
    """main4_NMF.py:99-103."""
    num = np.sum(np.asarray(ref, dtype=np.float64) ** 2)
    den = np.sum((np.asarray(ref, dtype=np.float64) - np.asarray(est, dtype=np.float64)) ** 2)
    return 10 * np.log10(num / (den + 1e-10))


# ---- fixture producers (generate_part{1,2}_data.py) ---------------------------------------------

def create_random_mask(n_samples, mask_ratio=0.3, max_gap_len=400):
    """generate_part1_data.py:26-35 (caller seeds np.random; the reference does not).
    True = keep, False = lost.  process_part1 calls it with mask_ratio=0.25 (:44)."""
    mask = np.ones(n_samples, dtype=bool)
    num_gaps = int(n_samples * mask_ratio / max_gap_len * 2)
    for _ in range(num_gaps):
        gap_len = np.random.randint(50, max_gap_len)
        gap_start = np.random.randint(0, n_samples - gap_len)
        mask[gap_start:gap_start + gap_len] = 0
    return mask


def centre_gap(n, sr, half_s=1.0):
    """generate_part2_data.py:36-40: the zeroed range [centre - sr, centre + sr)."""
    centre = n // 2
    half = int(half_s * sr)
    return centre - half, centre + half


# ---- sample-level detectors / baselines of the sibling scripts ------------------------------------
def find_main_gap(x, threshold=1e-4):
    """main3_AR_text_gap.py:34-49: (start, end) of the samples with |x| < threshold, or None."""
    idx = np.where(np.abs(x) < threshold)[0]
    if len(idx) == 0:
        return None
    return int(idx[0]), int(idx[-1] + 1)


def find_gaps(x, threshold=0.01, min_len=100):
    """main3_AR_text_mask.py:30-52: runs of |x| < threshold longer than min_len samples."""
    is_gap = (np.abs(x) < threshold)
    diff = np.diff(is_gap.astype(int))
    starts = np.where(diff == 1)[0] + 1
    ends = np.where(diff == -1)[0] + 1
    if is_gap[0]:
        starts = np.insert(starts, 0, 0)
    if is_gap[-1]:
        ends = np.append(ends, len(x))
    return [(int(s), int(e)) for s, e in zip(starts, ends) if (e - s) > min_len]


def linear_interp(x, threshold=1e-4):
    """linear_interp_part1.py:52-75: damaged = not(|x| > threshold) samples <- np.interp over the valid ones.
    Returns (y float32, number of damaged samples)."""
    mask = np.abs(x) > threshold
    n_damaged = int(np.sum(~mask))
    if n_damaged == 0 or n_damaged == len(x):
        return x.copy(), n_damaged
    x_all = np.arange(len(x))
    y = x.copy()
    y[~mask] = np.interp(x_all[~mask], x_all[mask], x[mask])
    return y, n_damaged
