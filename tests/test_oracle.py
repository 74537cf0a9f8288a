"""CPU tests that pin the oracle: restatement == library calls == reference outputs (golden)."""
import warnings

import numpy as np
import pytest
from scipy import signal

from oracle import libcalls, restate

SR = 44100


def rel_l2(a, b):
    a = np.asarray(a, np.complex128 if np.iscomplexobj(a) else np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def x_gap(golden):
    return libcalls.load_normalised(golden.gap_input_i16())


# ---- golden vectors -------------------------------------------------------------------------------

def test_c2_libcalls_reproduces_shipped_wav(golden, x_gap):
    """main4_NMF_gap.py end to end vs demo_assets/part2/fixed_nmf_gap.wav (<= 1 LSB) and vs the
    reference script's own float output run in the build container (y_sub: bit-equal)."""
    y, st = libcalls.restore_columns(x_gap, SR, return_all=True)
    c2 = golden.c2
    assert np.array_equal(st["bad"], c2["bad_cols"])
    assert st["bad"][0] == 690 and st["bad"][-1] == 1033 and len(st["bad"]) == 344
    assert st["n_iter"] == int(c2["n_iter"]) == 200
    assert abs(st["err"] - float(c2["err"])) <= 1e-6 * float(c2["err"])
    shipped = golden.gap_input_i16().astype(np.int32) + c2["shipped_minus_input_i16"]
    q = libcalls.quantise_int16(y).astype(np.int32)
    assert np.max(np.abs(q - shipped)) <= 1
    assert np.allclose(y[::32], c2["y_sub"], rtol=0, atol=2e-6)
    assert np.allclose(st["mag"][::16, ::16], c2["mag_sub"], rtol=1e-5, atol=1e-9)


def test_c3_libcalls_reproduces_reference_run(golden):
    x = libcalls.load_normalised(golden.mask_input_i16())
    y, st = libcalls.restore_columns(x, SR, threshold=0.01, frac=0.8, return_all=True)
    c3 = golden.c3
    assert np.array_equal(st["bad"], c3["bad_cols"]) and len(st["bad"]) == 288
    assert st["n_iter"] == int(c3["n_iter"])
    assert abs(st["err"] - float(c3["err"])) <= 1e-6 * float(c3["err"])
    ref = golden.mask_input_i16().astype(np.int32) + c3["ref_minus_input_i16"]
    assert np.max(np.abs(libcalls.quantise_int16(y).astype(np.int32) - ref)) <= 1


def test_c1_libcalls_reproduces_shipped_wavs(golden):
    c1 = golden.c1
    raw = c1["raw"]
    cor, gs, ge = libcalls.part0_apply_mask(raw, 0.2)
    assert (gs, ge) == tuple(c1["gap"]) == (882, 1323)
    assert np.array_equal(cor, c1["corrupted"])
    assert np.array_equal(libcalls.quantise_int16(raw), c1["shipped_original_i16"])
    assert np.array_equal(libcalls.quantise_int16(cor), c1["shipped_corrupted_i16"])
    y, st = libcalls.part0_restore(raw, cor, SR, gs, ge, return_all=True)
    assert st["cols"] == (6, 10)
    assert list(st["n_iters"]) == list(c1["n_iters"])
    q = libcalls.quantise_int16(y).astype(np.int32)
    assert np.max(np.abs(q - c1["shipped_restored_i16"].astype(np.int32))) <= 1
    assert libcalls.snr_db(c1["restored"], y) > 100


# ---- restatement vs the libraries -----------------------------------------------------------------

@pytest.mark.parametrize("N,n_fft,hop", [(2205, 512, 128), (441000, 1024, 256), (441000, 2048, 512),
                                         (5000, 256, 64), (4096, 512, 128), (1024, 1024, 256)])
def test_stft_restatement(N, n_fft, hop):
    rng = np.random.default_rng(N + n_fft)
    x = rng.standard_normal(N).astype(np.float32)
    _, _, Z = signal.stft(x, SR, nperseg=n_fft, noverlap=n_fft - hop)
    T, nadd = restate.stft_geometry(N, n_fft, hop)
    assert Z.shape == (n_fft // 2 + 1, T) and Z.dtype == np.complex64
    Zr = restate.stft(x, n_fft, hop)
    assert Zr.shape == (T, n_fft // 2 + 1)
    assert rel_l2(Zr, Z.T) < 1e-6
    assert rel_l2(np.abs(Zr), np.abs(Z.T)) < 1e-6


def test_stft_geometry_known():
    assert restate.stft_geometry(441000, 1024, 256) == (1724, 88)
    assert restate.stft_geometry(441000, 2048, 512) == (863, 344)
    assert restate.stft_geometry(2205, 512, 128) == (19, 99)
    assert restate.stft_geometry(158760000, 2048, 512)[0] == 310080


@pytest.mark.parametrize("N,n_fft,hop", [(2205, 512, 128), (44100, 1024, 256), (30000, 2048, 512)])
def test_istft_restatement(N, n_fft, hop):
    rng = np.random.default_rng(7)
    x = rng.standard_normal(N).astype(np.float32)
    _, _, Z = signal.stft(x, SR, nperseg=n_fft, noverlap=n_fft - hop)
    Z = (Z * (1 + 0.3 * rng.standard_normal(Z.shape))).astype(np.complex64)   # not a valid STFT any more
    _, y = signal.istft(Z, SR, nperseg=n_fft, noverlap=n_fft - hop)
    yr = restate.istft(np.ascontiguousarray(Z.T), n_fft, hop)
    assert y.dtype == np.float32 and y.shape == yr.shape
    assert rel_l2(yr, y) < 1e-6


def test_mask_fraction_is_integer_predicate():
    """np.mean(bool[ws:we]) > frac  <=>  den*cnt > num*len, exhaustively for len <= 4096."""
    for frac in (0.9, 0.8):
        num, den = restate.frac_to_ratio(frac)
        assert (num, den) in ((9, 10), (4, 5))
        for ln in range(1, 4097):
            cnt = np.arange(ln + 1)
            assert np.array_equal(cnt / ln > frac, den * cnt > num * ln), (frac, ln)


def test_mask_threshold_is_float32():
    assert np.float32(1e-4).view(np.uint32) == 0x38D1B717
    assert np.float32(0.01).view(np.uint32) == 0x3C23D70A
    x = np.array([np.float32(1e-4), np.nextafter(np.float32(1e-4), np.float32(0))], np.float32)
    assert list(np.abs(x) < 1e-4) == [False, True]


@pytest.mark.parametrize("thr,frac", [(1e-4, 0.9), (0.01, 0.8)])
def test_mask_restatement_random(thr, frac):
    rng = np.random.default_rng(3)
    num, den = restate.frac_to_ratio(frac)
    for N, n_fft, hop in [(20000, 1024, 256), (12345, 512, 128), (441000, 2048, 512), (700, 256, 64)]:
        x = rng.standard_normal(N).astype(np.float32) * 0.1
        for _ in range(12):
            a = rng.integers(0, N)
            x[a:a + rng.integers(1, 4 * hop)] = 0
        x[rng.integers(0, N, 50)] = np.float32(thr)          # exactly at the threshold: not a gap
        T, _ = restate.stft_geometry(N, n_fft, hop)
        want = libcalls.column_mask(x, T, hop, thr, frac)
        got = restate.column_mask(x, T, hop, thr, num, den)
        assert np.array_equal(want, got)


def test_mask_last_column_empty_window():
    """441000 @ 2048/512: centre of the last frame is 441344 > N -> empty window -> nan -> not bad."""
    x = np.zeros(441000, np.float32)
    T, _ = restate.stft_geometry(len(x), 2048, 512)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = libcalls.column_mask(x, T, 512, 1e-4, 0.9)
    got = restate.column_mask(x, T, 512, 1e-4, 9, 10)
    assert np.array_equal(want, got) and len(got) == T - 1


@pytest.mark.parametrize("seed", [0, 42, 123456789])
def test_rng_restatement_bit_equal(seed):
    want = np.random.RandomState(seed).standard_normal(5000)
    got = restate.standard_normal(seed, 5000)
    assert np.array_equal(want, got)


def test_init_factors_bit_equal_to_sklearn():
    from sklearn.decomposition._nmf import _initialize_nmf
    rng = np.random.default_rng(0)
    X = np.abs(rng.standard_normal((65, 50))).astype(np.float32)
    W, H = _initialize_nmf(X, 8, init="random", random_state=42)
    W0, Ht0 = restate.init_factors(X.mean(), 65, 50, 8, 42)
    assert np.array_equal(W, W0) and np.array_equal(H.T, Ht0)


@pytest.mark.parametrize("F,T,K,tol", [(65, 50, 8, 1e-4), (257, 19, 40, 1e-4), (129, 300, 16, 0.0)])
def test_cd_restatement_bit_equal_to_sklearn(F, T, K, tol):
    rng = np.random.default_rng(F)
    X = np.abs(rng.standard_normal((F, T))).astype(np.float32)
    W0, Ht0 = restate.init_factors(X.mean(), F, T, K, 0)
    W, H, n_iter, err = libcalls.nmf_fit(X, K, seed=0, max_iter=60, tol=tol)
    Wr, Htr, it, er = restate.nmf_cd(np.ascontiguousarray(X.T), W0, Ht0, 60, tol)
    assert it == n_iter
    assert np.array_equal(W, Wr) and np.array_equal(H.T, Htr)
    assert abs(er - err) <= 2e-6 * err


def test_mu_restatement_close_to_sklearn():
    rng = np.random.default_rng(5)
    X = np.abs(rng.standard_normal((40, 60))).astype(np.float32)
    W0, Ht0 = restate.init_factors(X.mean(), 40, 60, 6, 1)
    W, H, n_iter, err = libcalls.nmf_fit(X, 6, W0=W0, H0=Ht0.T, max_iter=30, tol=0.0, solver="mu")
    Wr, Htr, er = restate.nmf_mu_fro(np.ascontiguousarray(X.T), W0, Ht0, 30)
    assert n_iter == 30
    assert rel_l2(Wr, W) < 1e-4 and abs(er - err) < 1e-4 * err


def test_restate_pipeline_matches_libcalls_c2(golden, x_gap):
    y, st = restate.restore_columns(x_gap)
    yl, sl = libcalls.restore_columns(x_gap, SR, return_all=True)
    assert np.array_equal(st["bad"], sl["bad"])
    assert rel_l2(st["V"], sl["mag"].T) < 1e-6
    assert np.allclose(st["fill"], sl["X"][:, sl["bad"][0]], rtol=1e-5)
    assert st["n_iter"] == sl["n_iter"]
    assert abs(st["err"] - sl["err"]) < 1e-4 * sl["err"]
    gs, ge = golden.c2["gap"]
    assert libcalls.snr_db(yl, y) > 60 and libcalls.snr_db(yl[gs:ge], y[gs:ge]) > 40


def test_f4_sibling_detectors_reproduce_reference_scripts(golden):
    """oracle restatements vs outputs of the unmodified main3_AR_text_gap.py / main3_AR_text_mask.py class
    methods and of linear_interp_part1.py run as shipped (tests/golden/make_golden.py --siblings)."""
    f4 = golden.f4
    x_gap = libcalls.load_normalised(golden.gap_input_i16())
    assert libcalls.find_main_gap(x_gap, 1e-4) == tuple(int(v) for v in f4["main_gap"])
    assert libcalls.find_gaps(x_gap, 0.01, 100) == [tuple(int(v) for v in r) for r in f4["gaps_on_gap"]]
    dr = f4["damaged_random_i16"]
    assert libcalls.find_gaps(libcalls.load_normalised(dr), 0.01, 100) == [tuple(int(v) for v in r) for r in f4["gaps_random"]]
    xr = dr.astype(np.float32) / np.max(np.abs(dr))                      # linear_interp_part1.py:46
    y, nd = libcalls.linear_interp(xr, 1e-4)
    assert nd == int(f4["n_damaged"])
    assert np.array_equal(libcalls.quantise_int16(y), f4["fixed_linear_i16"])
