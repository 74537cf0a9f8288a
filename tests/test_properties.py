"""Property tests (hypothesis) of the byte/index contracts on the CPU side: the integer mask
predicate the CUDA kernel implements against the reference's float comparison, the STFT framing against scipy's own
shapes, and the analysis/synthesis round trip.  The same predicate and framing run on the device in tests/test_gpu_parity.py;
here hypothesis searches the parameter space (ragged lengths, gaps at the edges, thresholds hit exactly)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import libcalls, restate

GEOMS = [(64, 16), (128, 32), (256, 64), (512, 128), (1024, 256), (2048, 512), (256, 256), (512, 256)]


@st.composite
def gapped_signal(draw):
    n_fft, hop = draw(st.sampled_from(GEOMS))
    N = draw(st.integers(min_value=n_fft, max_value=6 * n_fft + 777))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    thr, frac = draw(st.sampled_from([(1e-4, 0.9), (0.01, 0.8)]))
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(N) * 0.1).astype(np.float32)
    for _ in range(draw(st.integers(0, 6))):
        a = draw(st.integers(0, N - 1))
        ln = draw(st.integers(1, 3 * hop))
        x[a:a + ln] = 0
    k = draw(st.integers(0, 40))                      # samples that sit exactly on / one ulp below the threshold
    x[rng.integers(0, N, k)] = np.float32(thr)
    x[rng.integers(0, N, k)] = np.nextafter(np.float32(thr), np.float32(0))
    if draw(st.booleans()):
        x[:draw(st.integers(1, hop))] = 0             # a gap that touches the left edge
    if draw(st.booleans()):
        x[N - draw(st.integers(1, hop)):] = 0         # ... and the right edge (the last window may be empty)
    return x, n_fft, hop, thr, frac


@settings(max_examples=60, deadline=None)
@given(gapped_signal())
def test_integer_mask_predicate_equals_the_reference_comparison(case):
    x, n_fft, hop, thr, frac = case
    T, _ = restate.stft_geometry(len(x), n_fft, hop)
    num, den = restate.frac_to_ratio(frac)
    want = libcalls.column_mask(x, T, hop, thr, frac)                 # np.mean(bool) > frac, as the reference writes it
    got = restate.column_mask(x, T, hop, thr, num, den)               # den * cnt > num * len, what mask.cu computes
    assert np.array_equal(got, want)


@settings(max_examples=40, deadline=None)
@given(st.sampled_from(GEOMS), st.integers(min_value=0, max_value=5000), st.integers(0, 2 ** 31 - 1))
def test_stft_framing_matches_scipy_and_round_trips(geom, extra, seed):
    from scipy import signal
    n_fft, hop = geom
    N = n_fft + extra
    x = np.random.default_rng(seed).standard_normal(N).astype(np.float32)
    T, nadd = restate.stft_geometry(N, n_fft, hop)
    _, t, Zs = signal.stft(x, 44100, nperseg=n_fft, noverlap=n_fft - hop)
    assert Zs.shape == (n_fft // 2 + 1, T) and (N + nadd) % hop == 0 and 0 <= nadd < n_fft
    Z = restate.stft(x, n_fft, hop)
    assert np.linalg.norm(Z.T - Zs) <= 1e-5 * np.linalg.norm(Zs)     # north_star: 1e-5 relative L2
    if n_fft // hop >= 2:                                             # NOLA holds for Hann with >= 50 % overlap
        y = restate.istft(Z, n_fft, hop, N)
        assert np.linalg.norm(y - x) <= 1e-5 * np.linalg.norm(x)


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 40), st.integers(2, 30), st.integers(1, 8), st.integers(0, 2 ** 31 - 1))
def test_cd_sweep_never_increases_the_objective_and_keeps_factors_nonnegative(F, T, K, seed):
    """Two invariants of the coordinate-descent iteration the kernels implement (_cdnmf_fast.pyx:8-38)."""
    rng = np.random.default_rng(seed)
    X = np.abs(rng.standard_normal((F, T))).astype(np.float32)
    W0, Ht0 = restate.init_factors(float(X.mean()), F, T, K, seed % 1000)
    obj = lambda W, Ht: float(np.linalg.norm(X.astype(np.float64) - W.astype(np.float64) @ Ht.astype(np.float64).T))
    prev = obj(W0, Ht0)
    W, Ht = W0.copy(), Ht0.copy()
    for _ in range(4):
        restate.cd_sweep(W, (Ht.T @ Ht).astype(np.float32), (X @ Ht).astype(np.float32))
        restate.cd_sweep(Ht, (W.T @ W).astype(np.float32), (X.T @ W).astype(np.float32))
        cur = obj(W, Ht)
        assert cur <= prev * (1 + 1e-5) + 1e-6 * float(np.linalg.norm(X)) and W.min() >= 0 and Ht.min() >= 0   # (an exact fit sits at rounding noise)
        prev = cur
