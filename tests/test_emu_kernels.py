"""Kernel-logic tests without a GPU: the CUDA sources compiled for the host against csrc/emu/cuda_emu.h
(tests/emu_harness.py) and driven through the same C ABI, compared with the oracle on small seeded inputs.
This checks indexing, tiling, reductions and the host-side orchestration; the real parity tests are the
`-m gpu` ones, which run the sm_100a build."""
import numpy as np
import pytest
from scipy import signal

import emu_harness as E
from oracle import libcalls, restate

SR = 8000


def rel_l2(a, b):
    return float(np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300))


@pytest.mark.parametrize("N,n_fft,hop", [(3000, 128, 32), (1111, 64, 16), (2500, 256, 64)])
def test_emu_stft_istft(N, n_fft, hop):
    rng = np.random.default_rng(N)
    x = rng.standard_normal((2, N)).astype(np.float32)
    mag, Z = E.stft(x, n_fft, hop)
    for b in range(2):
        _, _, Zs = signal.stft(x[b], SR, nperseg=n_fft, noverlap=n_fft - hop)
        assert Z[b].shape == Zs.shape
        assert rel_l2(Z[b], Zs) < 1e-6 and rel_l2(mag[b], np.abs(Zs)) < 1e-6
    _, _, Zs = signal.stft(x[0], SR, nperseg=n_fft, noverlap=n_fft - hop)
    Zs = (Zs * (1 + 0.2 * rng.standard_normal(Zs.shape))).astype(np.complex64)
    _, ys = signal.istft(Zs, SR, nperseg=n_fft, noverlap=n_fft - hop)
    y = E.istft(Zs, n_fft, hop, N)
    assert rel_l2(y[0], ys[:N]) < 1e-6


@pytest.mark.parametrize("thr,num,den,frac", [(1e-4, 9, 10, 0.9), (0.01, 4, 5, 0.8)])
def test_emu_mask_bit_exact(thr, num, den, frac):
    rng = np.random.default_rng(4)
    for N, n_fft, hop in [(5000, 128, 32), (3333, 256, 64), (900, 64, 16)]:
        x = rng.standard_normal((2, N)).astype(np.float32) * 0.1
        for b in range(2):
            for _ in range(8):
                a = rng.integers(0, N)
                x[b, a:a + rng.integers(1, 5 * hop)] = 0
            x[b, rng.integers(0, N, 20)] = np.float32(thr)
        T, _ = restate.stft_geometry(N, n_fft, hop)
        bad, idx, nb = E.gap_mask(x, hop, T, thr, num, den)
        for b in range(2):
            want = libcalls.column_mask(x[b], T, hop, thr, frac)
            assert nb[b] == len(want) and np.array_equal(idx[b, :nb[b]], want)


@pytest.mark.parametrize("F,T,K,iters", [(65, 95, 8, 4), (33, 70, 40, 3), (130, 40, 100, 2)])
def test_emu_nmf_fit(F, T, K, iters):
    rng = np.random.default_rng(F)
    X = np.abs(rng.standard_normal((F, T))).astype(np.float32)
    W, H, err, nit = E.nmf_fit(X, K, max_iter=iters, tol=0.0, seed=42)
    Wo, Ho, no, eo = libcalls.nmf_fit(X, K, seed=42, max_iter=iters, tol=0.0)
    assert nit[0] == no
    assert abs(err[0] - eo) < 1e-5 * eo
    assert rel_l2(W[0], Wo) < 1e-4 and rel_l2(H[0], Ho) < 1e-4


@pytest.mark.parametrize("F,T,K,B,iters", [(129, 50, 40, 4, 3), (140, 36, 20, 8, 3), (258, 24, 40, 4, 2), (261, 20, 24, 4, 2)])
def test_emu_nmf_fit_batch_one_thread_per_row_w_side(F, T, K, B, iters):
    """A batch large enough (for the emulator's 4 SMs) that the W side takes its one-thread-per-row shape: the blocked
    sweep (blocks of 8 coordinates, in-block corrections) must reproduce sklearn's sequential sweep; F = 129 leaves the
    last block of every clip with a single row (as F = 513 does at the real sizes); F = 258 / 261 are 256 rows plus a
    remainder that rides with the last block on its extra warp (nmf_wside.cu: TAIL)."""
    rng = np.random.default_rng(F + B)
    X = np.abs(rng.standard_normal((B, F, T))).astype(np.float32)
    X[1, :, 5:9] = 0.0                                     # a clip with zero columns: some coordinates hit the bound
    W, H, err, nit = E.nmf_fit(X, K, max_iter=iters, tol=0.0, seed=42)
    for b in range(B):
        Wo, Ho, no, eo = libcalls.nmf_fit(X[b], K, seed=42, max_iter=iters, tol=0.0)
        assert nit[b] == no
        assert abs(err[b] - eo) < 1e-5 * eo
        assert rel_l2(W[b], Wo) < 1e-4 and rel_l2(H[b], Ho) < 1e-4


def test_emu_whole_path_and_edge_cases():
    rng = np.random.default_rng(2)
    N, n_fft, hop, K = 6000, 128, 32, 8
    t = np.arange(N) / SR
    x = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.3 * np.sin(2 * np.pi * 1230 * t) + 0.05 * rng.standard_normal(N)).astype(np.float32)
    x[2500:3300] = 0
    x /= np.abs(x).max()
    x2 = x.copy()
    x2[2500:3300] = x[1000:1800]                      # a clip with nothing to restore
    r = E.inpaint(np.stack([x, x2]), n_fft=n_fft, hop=hop, rank=K, max_iter=5, tol=1e-4, seed=42)
    y, st = libcalls.restore_columns(x, SR, n_fft=n_fft, hop=hop, K=K, seed=42, max_iter=5, return_all=True)
    assert np.array_equal(r["bad_idx"][0][:r["n_bad"][0]], st["bad"])
    assert r["n_iter"][0] == st["n_iter"] and abs(r["err"][0] - st["err"]) < 1e-5 * st["err"]
    assert libcalls.snr_db(y, r["y"][0]) > 90 and libcalls.snr_db(y[2500:3300], r["y"][0][2500:3300]) > 70
    assert r["n_bad"][1] == 0 and np.array_equal(r["y"][1], x2)      # input returned untouched
    with pytest.raises(E.capi.AinmfError) as e:
        E.inpaint(np.zeros(6016, np.float32), n_fft=n_fft, hop=hop, rank=K, max_iter=2)   # 6016 = 188*hop: no empty window
    assert e.value.code == E.capi.ERR_ALL_BAD


def test_emu_part0_variant():
    """col_start/col_end + n_outer (main4_NMF.py:74-90) against the library-call oracle."""
    rng = np.random.default_rng(6)
    sr, N, n_fft, hop, K = 8000, 1600, 128, 32, 8
    t = np.arange(N) / sr
    raw = (0.6 * np.sin(2 * np.pi * 300 * t) + 0.2 * np.sin(2 * np.pi * 900 * t) + 0.02 * rng.standard_normal(N)).astype(np.float32)
    cor, gs, ge = libcalls.part0_apply_mask(raw, 0.2)
    yo, so = libcalls.part0_restore(raw, cor, sr, gs, ge, n_fft=n_fft, hop=hop, K=K, n_outer=2, seed=0,
                                    max_iter=12, return_all=True)
    cs, ce = so["cols"]
    r = E.inpaint(cor, n_fft=n_fft, hop=hop, rank=K, max_iter=12, tol=1e-4, seed=0, col_start=cs, col_end=ce, n_outer=2)
    assert abs(int(r["n_iter"][0]) - so["n_iters"][-1]) <= 1
    assert libcalls.snr_db(so["pre_blend"], r["y"][0]) > 60


def test_emu_invalid_arguments():
    x = np.zeros(4000, np.float32)
    for kw in (dict(n_fft=100, hop=25), dict(n_fft=128, hop=48), dict(rank=0), dict(rank=129), dict(max_iter=0),
               dict(n_fft=8192, hop=2048), dict(col_start=0, col_end=3), dict(n_outer=0), dict(solver=7)):
        p = dict(n_fft=128, hop=32, rank=8, max_iter=2)
        p.update(kw)
        with pytest.raises(E.capi.AinmfError) as e:
            E.inpaint(x, **p)
        assert e.value.code == E.capi.ERR_INVALID, kw
    with pytest.raises(E.capi.AinmfError):
        E.stft(np.zeros(50, np.float32), 128, 32)


@pytest.mark.parametrize("KP,rows", [(32, 200), (64, 130), (128, 257)])
def test_emu_incremental_sweep_matches_reference_sweep(KP, rows):
    """cd_sweep_rows_inc (the sweep of the tensor-core h-step) against the reference sweep order of
    _cdnmf_fast.pyx, on random non-negative factors with some exact zeros."""
    import ctypes as C
    rng = np.random.default_rng(KP)
    A = np.abs(rng.standard_normal((rows, KP))).astype(np.float32)
    A[rng.random((rows, KP)) < 0.2] = 0
    Hm = np.abs(rng.standard_normal((KP + 40, KP))).astype(np.float32)
    G = (Hm.T @ Hm).astype(np.float32)
    G[:, 5] = 0; G[5, :] = 0                                   # a dead component: zero Gram diagonal -> skipped
    Bm = (np.abs(rng.standard_normal((rows, KP))) * KP).astype(np.float32)
    want = A.copy()
    v_want = restate.cd_sweep(want, G, Bm)
    got = A.copy()
    viol = np.zeros((rows + 127) // 128, np.float32)
    fn = E.lib().ainmf_test_sweep_inc
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    assert fn(E.ptr(got), E.ptr(G), E.ptr(Bm), rows, KP, E.ptr(viol), None) == 0
    scale = np.abs(want).max()
    assert np.abs(got - want).max() < 2e-4 * scale
    assert abs(viol.sum() - v_want) < 1e-3 * v_want


@pytest.mark.parametrize("F,T,K,iters,tol", [(65, 95, 8, 12, 0.0), (40, 60, 6, 30, 1e-3)])
def test_emu_mu_solver_matches_sklearn_mu(F, T, K, iters, tol):
    """solver='mu' (Frobenius multiplicative update) against sklearn's, same initial factors; with tol > 0 the
    every-10-iterations stop test must fire at the same iteration."""
    rng = np.random.default_rng(F + K)
    X = np.abs(rng.standard_normal((F, T))).astype(np.float32)
    W0, Ht0 = restate.init_factors(X.mean(), F, T, K, 3)
    Wo, Ho, no, eo = libcalls.nmf_fit(X, K, W0=W0, H0=Ht0.T, max_iter=iters, tol=tol, solver="mu")
    W, H, err, nit = E.nmf_fit(X, K, max_iter=iters, tol=tol, W0=W0, H0=np.ascontiguousarray(Ht0.T), solver=E.capi.SOLVER_MU)
    assert nit[0] == no
    assert abs(err[0] - eo) < 1e-4 * eo
    assert rel_l2(W[0], Wo) < 1e-3 and rel_l2(H[0], Ho) < 1e-3


@pytest.mark.parametrize("F,T,K,iters,tol", [(65, 95, 8, 12, 0.0), (70, 130, 40, 6, 0.0), (40, 60, 6, 40, 1e-3)])
def test_emu_mu_kl_solver_matches_sklearn(F, T, K, iters, tol):
    """solver='mu-kl' (fused ratio kernels, nmf_mukl.cu) against sklearn solver='mu', beta_loss='kullback-leibler' from
    the same initial factors: factors, n_iter_ (every-10th-iteration test) and reconstruction_err_ = sqrt(2 D_KL).
    Zero columns exercise the x <= eps and W.H < eps branches."""
    rng = np.random.default_rng(F + K)
    X = np.abs(rng.standard_normal((F, T))).astype(np.float32)
    X[:, 7:11] = 0.0
    X[3, :] = 0.0
    W0, Ht0 = restate.init_factors(X.mean(), F, T, K, 3)
    Wo, Ho, no, eo = libcalls.nmf_fit(X, K, W0=W0, H0=Ht0.T, max_iter=iters, tol=tol, solver="mu", beta_loss="kullback-leibler")
    W, H, err, nit = E.nmf_fit(X, K, max_iter=iters, tol=tol, W0=W0, H0=np.ascontiguousarray(Ht0.T), solver=E.capi.SOLVER_MU_KL)
    assert nit[0] == no
    assert abs(err[0] - eo) < 1e-4 * eo
    assert rel_l2(W[0], Wo) < 1e-3 and rel_l2(H[0], Ho) < 1e-3


# ---- callers / baselines either side of the NMF path ---------------------------------------
def _gap_signals():
    rng = np.random.default_rng(7)
    out = []
    for N, kind in ((5000, "gaps"), (9000, "edges"), (3000, "none"), (2049, "all")):
        x = (rng.standard_normal(N) * 0.3).astype(np.float32)
        if kind == "gaps":
            for _ in range(12):
                s = int(rng.integers(0, N - 500)); l = int(rng.integers(1, 450))
                x[s:s + l] = 0
            x[rng.integers(0, N, 50)] = 5e-5
        elif kind == "edges":
            x[:300] = 0; x[-1234:] = 0; x[3000:3101] = 0; x[3500:3600] = 0; x[4096:6144] = 0
        elif kind == "all":
            x[:] = 0
        out.append(x)
    return out


@pytest.mark.parametrize("thr", [1e-4, 0.01])
def test_emu_gap_detectors_bit_exact(thr):
    for x in _gap_signals():
        span = E.find_main_gap(x[None], thr)[0]
        ref = libcalls.find_main_gap(x, thr)
        assert tuple(span) == (ref if ref is not None else (-1, -1))
        runs, n = E.find_gaps(x[None], thr, 100, 64)
        ref_runs = libcalls.find_gaps(x, thr, 100)
        assert int(n[0]) == len(ref_runs)
        assert [tuple(r) for r in runs[0, :len(ref_runs)]] == ref_runs


def test_emu_linear_interp_blend_snr():
    for x in _gap_signals():
        y, nd = E.linear_interp(x[None], 1e-4)
        yo, no = libcalls.linear_interp(x, 1e-4)
        assert int(nd[0]) == no
        assert np.max(np.abs(y[0].astype(np.float64) - yo.astype(np.float64))) <= 1.2e-7 * max(1.0, float(np.max(np.abs(yo))))
    rng = np.random.default_rng(3)
    raw = (rng.standard_normal(2205) * 0.4).astype(np.float32)
    res = (raw + 0.05 * rng.standard_normal(2205)).astype(np.float32)
    out = E.blend_boundaries(raw, res, 882, 1323, 50)
    ref = libcalls.part0_blend(raw, res, 882, 1323)
    assert np.array_equal(out, ref)
    assert abs(E.snr_db(raw, out, 882, 1323) - libcalls.snr_db(raw[882:1323], ref[882:1323])) < 1e-3


def test_emu_apply_gaps_matches_generate_part1():
    """create_random_mask semantics (generate_part1_data.py:26-35, 44-46): gap list drawn on the host, zeroed on the device."""
    rng = np.random.default_rng(5)
    N = 20000
    x = (rng.standard_normal((3, N)) * 0.3 + 1.0).astype(np.float32)
    starts = np.empty((3, 40), np.int64); lens = np.empty((3, 40), np.int64)
    for b in range(3):
        np.random.seed(b)
        for g in range(40):
            l = np.random.randint(50, 400); s = np.random.randint(0, N - l)
            starts[b, g], lens[b, g] = s, l
    starts[2, 5], lens[2, 5] = N - 10, 400          # clipped at N
    lens[1, 7] = 0                                  # skipped entry
    starts[0, 3] = -1                               # skipped entry
    ref = x.copy()
    for b in range(3):
        for g in range(40):
            if starts[b, g] >= 0 and lens[b, g] > 0:
                ref[b, starts[b, g]:min(N, starts[b, g] + lens[b, g])] = 0
    out = E.apply_gaps(x, starts, lens)
    assert np.array_equal(out, ref)
