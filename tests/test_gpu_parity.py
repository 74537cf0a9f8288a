"""GPU parity tests: the CUDA path (through torch.ops.ainmf -> C ABI -> sm_100a kernels) against the oracle.

Tolerances are north_star's: masks / indices bit-exact; |Z| rel-L2 <= 1e-5; objective rel <= 1e-4;
restored-waveform SNR vs the reference output >= 60 dB (in-gap SNR is gated at >= 60 dB as well, the
whole-wave figure being dominated by untouched frames).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import libcalls, restate  # noqa: E402

SR = 44100


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    import ainmf
    return ainmf.ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- K1 / K5 --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,n_fft,hop", [(2205, 512, 128), (441000, 1024, 256), (441000, 2048, 512),
                                         (5000, 256, 64), (4096, 64, 16), (100000, 4096, 1024), (1024, 1024, 256)])
def test_stft_matches_scipy(ops, N, n_fft, hop):
    from scipy import signal
    rng = np.random.default_rng(N)
    x = rng.standard_normal((2, N)).astype(np.float32)
    mag, Z = ops.stft(dev(x), n_fft, hop)
    for b in range(2):
        _, _, Zs = signal.stft(x[b], SR, nperseg=n_fft, noverlap=n_fft - hop)
        assert tuple(Z[b].shape) == Zs.shape
        assert rel_l2(mag[b].cpu().numpy(), np.abs(Zs)) <= 1e-5          # north_star: 1e-5 relative L2
        assert rel_l2(Z[b].cpu().numpy(), Zs) <= 1e-5


@pytest.mark.parametrize("N,n_fft,hop", [(2205, 512, 128), (441000, 1024, 256), (30000, 2048, 512), (4096, 64, 16)])
def test_istft_matches_scipy(ops, N, n_fft, hop):
    from scipy import signal
    rng = np.random.default_rng(7)
    x = rng.standard_normal(N).astype(np.float32)
    _, _, Z = signal.stft(x, SR, nperseg=n_fft, noverlap=n_fft - hop)
    Z = (Z * (1 + 0.3 * rng.standard_normal(Z.shape))).astype(np.complex64)
    _, ys = signal.istft(Z, SR, nperseg=n_fft, noverlap=n_fft - hop)
    y = ops.istft(dev(Z[None]), n_fft, hop, N)[0].cpu().numpy()
    assert rel_l2(y, ys[:N]) <= 1e-5


def test_window_argument_matches_scipy(ops):
    """north_star lists `window` among the arguments of the path: a caller-supplied window (here Hamming and a Tukey-shaped
    array) must give scipy's stft / istft with `window=array`; None restores the periodic Hann default."""
    from scipy import signal
    rng = np.random.default_rng(4)
    n_fft, hop, N = 512, 128, 20000
    x = rng.standard_normal(N).astype(np.float32)
    try:
        for w in (signal.get_window("hamming", n_fft), signal.get_window(("tukey", 0.7), n_fft) + 0.05):
            ops.set_window(n_fft, w)
            mag, Z = ops.stft(dev(x[None]), n_fft, hop)
            _, _, Zs = signal.stft(x, SR, window=w.astype(np.float32), nperseg=n_fft, noverlap=n_fft - hop)
            assert rel_l2(Z[0].cpu().numpy(), Zs) <= 1e-5
            Zm = (Zs * (1 + 0.2 * rng.standard_normal(Zs.shape))).astype(np.complex64)
            _, ys = signal.istft(Zm, SR, window=w.astype(np.float32), nperseg=n_fft, noverlap=n_fft - hop)
            y = ops.istft(dev(Zm[None]), n_fft, hop, N)[0].cpu().numpy()
            assert rel_l2(y, ys[:N]) <= 1e-5
    finally:
        ops.set_window(n_fft, None)
    _, Z = ops.stft(dev(x[None]), n_fft, hop)
    _, _, Zs = signal.stft(x, SR, nperseg=n_fft, noverlap=n_fft - hop)
    assert rel_l2(Z[0].cpu().numpy(), Zs) <= 1e-5


def test_stft_istft_round_trip_full_size(ops):
    """Size-independent property at BASELINE sizes: istft(stft(x)) == x (NOLA holds for Hann, 75 % overlap)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((8, 441000), generator=g, device="cuda")
    for n_fft, hop in ((1024, 256), (2048, 512)):
        _, Z = ops.stft(x, n_fft, hop)
        y = ops.istft(Z, n_fft, hop, x.shape[1])
        err = (y - x).norm() / x.norm()
        assert float(err) < 1e-5


def test_stft_is_linear(ops):
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randn((1, 50000), generator=g, device="cuda")
    b = torch.randn((1, 50000), generator=g, device="cuda")
    _, Za = ops.stft(a, 1024, 256)
    _, Zb = ops.stft(b, 1024, 256)
    _, Zc = ops.stft(2 * a - 3 * b, 1024, 256)
    assert float((Zc - (2 * Za - 3 * Zb)).abs().max() / Zc.abs().max()) < 1e-5


# ---- K2 (bit-exact) --------------------------------------------------------------------------------------
@pytest.mark.parametrize("thr,num,den,frac", [(1e-4, 9, 10, 0.9), (0.01, 4, 5, 0.8)])
def test_mask_bit_exact_random(ops, thr, num, den, frac):
    rng = np.random.default_rng(3)
    for N, n_fft, hop in [(20000, 1024, 256), (12345, 512, 128), (441000, 2048, 512), (700, 256, 64), (441000, 1024, 256)]:
        x = rng.standard_normal(N).astype(np.float32) * 0.1
        for _ in range(20):
            a = rng.integers(0, N)
            x[a:a + rng.integers(1, 4 * hop)] = 0
        x[rng.integers(0, N, 50)] = np.float32(thr)
        x[rng.integers(0, N, 50)] = np.nextafter(np.float32(thr), np.float32(0))
        T, _ = restate.stft_geometry(N, n_fft, hop)
        want = libcalls.column_mask(x, T, hop, thr, frac)
        bad, idx, nb = ops.gap_mask(dev(x[None]), hop, T, thr, num, den)
        n = int(nb[0])
        assert n == len(want)
        assert np.array_equal(idx[0, :n].cpu().numpy().astype(np.int64), want)
        assert np.array_equal(np.nonzero(bad[0].cpu().numpy())[0], want)


def test_mask_golden_c2_c3(ops, golden):
    x2 = libcalls.load_normalised(golden.gap_input_i16())
    _, idx, nb = ops.gap_mask(dev(x2[None]), 256, 1724, 1e-4, 9, 10)
    assert np.array_equal(idx[0, :int(nb[0])].cpu().numpy(), golden.c2["bad_cols"])
    x3 = libcalls.load_normalised(golden.mask_input_i16())
    _, idx, nb = ops.gap_mask(dev(x3[None]), 256, 1724, 0.01, 4, 5)
    assert np.array_equal(idx[0, :int(nb[0])].cpu().numpy(), golden.c3["bad_cols"])


def test_mask_edge_cases(ops):
    # all silent: every column with a non-empty window is bad; 441000 @ 2048/512 has an empty last window
    x = np.zeros((1, 441000), np.float32)
    T, _ = restate.stft_geometry(441000, 2048, 512)
    _, idx, nb = ops.gap_mask(dev(x), 512, T, 1e-4, 9, 10)
    assert int(nb[0]) == T - 1 and int(idx[0, T - 2]) == T - 2
    # nothing silent
    _, _, nb = ops.gap_mask(dev(np.ones((3, 5000), np.float32)), 256, 21, 1e-4, 9, 10)
    assert nb.cpu().tolist() == [0, 0, 0]


# ---- front / back end (bit-exact) ------------------------------------------------------------------------
def test_pcm_load_store_bit_exact(ops, golden):
    pcm = golden.gap_input_i16()
    x, peak = ops.load_pcm16(dev(pcm[None]))
    want = libcalls.load_normalised(pcm)
    assert np.array_equal(x[0].cpu().numpy(), want)
    assert float(peak[0]) == float(np.max(np.abs(pcm)))
    rng = np.random.default_rng(0)
    st = rng.integers(-32768, 32767, (2, 3000, 2)).astype(np.int16)
    xs, _ = ops.load_pcm16(dev(st))
    for b in range(2):
        assert np.array_equal(xs[b].cpu().numpy(), libcalls.load_normalised(st[b]))
    y = (rng.standard_normal(10000) * 0.6).astype(np.float32)
    assert np.array_equal(ops.store_pcm16(dev(y)).cpu().numpy(), libcalls.quantise_int16(y))


# ---- K4 --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F,T,K,iters", [(65, 95, 8, 20), (257, 19, 40, 30), (513, 300, 64, 15), (129, 700, 128, 10),
                                         (1025, 90, 33, 10)])
def test_nmf_fit_matches_sklearn_custom_init(ops, F, T, K, iters):
    rng = np.random.default_rng(F + T)
    X = np.abs(rng.standard_normal((F, T))).astype(np.float32)
    W0, Ht0 = restate.init_factors(X.mean(), F, T, K, 0)
    Wo, Ho, no, eo = libcalls.nmf_fit(X, K, W0=W0, H0=Ht0.T, max_iter=iters, tol=0.0)
    W, H, err, nit = ops.nmf_fit(dev(X[None]), K, iters, 0.0, 0, dev(W0[None]), dev(np.ascontiguousarray(Ht0.T)[None]))
    assert int(nit[0]) == no == iters
    assert abs(float(err[0]) - eo) <= 1e-4 * eo                       # north_star: objective within 1e-4 relative
    assert rel_l2(W[0].cpu().numpy(), Wo) < 1e-3 and rel_l2(H[0].cpu().numpy(), Ho) < 1e-3


@pytest.mark.parametrize("B,F,T,K,iters", [(160, 257, 130, 64, 6), (40, 257, 200, 64, 6), (12, 513, 140, 40, 6), (3, 200, 260, 128, 5),
                                           (70, 129, 150, 20, 6), (60, 257, 130, 64, 5)])
def test_nmf_fit_batch_shapes_of_the_w_side_kernel(ops, B, F, T, K, iters):
    """The W-side kernel picks lanes per row / rows per block from the batch size (one thread per row for big batches, up
    to 8 lanes in 32-row blocks for a single clip); every shape must agree with sklearn on the same inputs."""
    rng = np.random.default_rng(B * 1000 + F)
    X = np.abs(rng.standard_normal((B, F, T))).astype(np.float32)
    W, H, err, nit = ops.nmf_fit(dev(X), K, iters, 0.0, 3, None, None)
    for b in sorted({0, B // 2, B - 1}):
        Wo, Ho, no, eo = libcalls.nmf_fit(X[b], K, seed=3, max_iter=iters, tol=0.0)
        assert int(nit[b]) == no == iters
        assert abs(float(err[b]) - eo) <= 1e-4 * eo
        assert rel_l2(W[b].cpu().numpy(), Wo) < 1e-3 and rel_l2(H[b].cpu().numpy(), Ho) < 1e-3


def test_nmf_fit_seeded_init_and_early_stop(ops):
    """init='random' with random_state: same initial factors as sklearn, same stop iteration (+-1)."""
    rng = np.random.default_rng(11)
    F, T, K = 257, 19, 40
    base = np.abs(rng.standard_normal((F, 6))).astype(np.float32) @ np.abs(rng.standard_normal((6, T))).astype(np.float32)
    X = (base + 0.01 * np.abs(rng.standard_normal((F, T)))).astype(np.float32)
    Wo, Ho, no, eo = libcalls.nmf_fit(X, K, seed=0, max_iter=200, tol=1e-4)
    W, H, err, nit = ops.nmf_fit(dev(X[None]), K, 200, 1e-4, 0, None, None)
    assert no < 200, "oracle should stop early on this matrix"
    assert int(nit[0]) == no                                          # small problem: reference-order violation sum
    assert abs(float(err[0]) - eo) <= 1e-4 * eo
    # one iteration from the seeded init must agree tightly
    Wo1, Ho1, _, _ = libcalls.nmf_fit(X, K, seed=0, max_iter=1, tol=0.0)
    W1, H1, _, _ = ops.nmf_fit(dev(X[None]), K, 1, 0.0, 0, None, None)
    assert rel_l2(W1[0].cpu().numpy(), Wo1) < 1e-5 and rel_l2(H1[0].cpu().numpy(), Ho1) < 1e-5


@pytest.mark.parametrize("seed,n", [(0, 7), (42, 100003), (0, 3000000), (123456789, 624 * 7 + 1)])
def test_device_rng_is_numpy_randomstate(ops, seed, n):
    """The initial factors come from RandomState(seed).standard_normal: MT19937 + the legacy polar method with its cached
    second deviate.  The device generator must reproduce numpy's float32 values -- bit for bit except where CUDA's and
    glibc's double log differ in the last bit AND that bit decides a float32 rounding (probability ~1e-9 per value)."""
    want = np.random.RandomState(seed).standard_normal(n).astype(np.float32)
    got = ops.standard_normal(seed, n).cpu().numpy()
    assert got.shape == want.shape
    same = got.view(np.uint32) == want.view(np.uint32)
    assert same.mean() >= 1 - 2e-6, (int((~same).sum()), n)
    assert np.max(np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))) <= 1      # at most one float32 ulp


def test_nmf_fit_batch_is_independent(ops):
    rng = np.random.default_rng(5)
    X = np.abs(rng.standard_normal((3, 129, 200))).astype(np.float32)
    W, H, err, nit = ops.nmf_fit(dev(X), 16, 12, 0.0, 42, None, None)
    for b in range(3):
        Wb, Hb, eb, nb = ops.nmf_fit(dev(X[b:b + 1]), 16, 12, 0.0, 42, None, None)
        assert torch.equal(W[b], Wb[0]) and torch.equal(H[b], Hb[0])      # deterministic reductions
        assert float(err[b]) == float(eb[0])


def test_tensor_core_path_skips_converged_clips(ops):
    """Persistent tensor-core kernels walk (clip, tile) lists and skip clips whose stop rule fired: a batch in which one
    clip converges early must give every clip exactly what it gets alone (K = 64, T >= 128: tcgen05 path)."""
    rng = np.random.default_rng(21)
    F, T, K = 257, 300, 64
    easy = np.abs(rng.standard_normal((F, 3))).astype(np.float32) @ np.abs(rng.standard_normal((3, T))).astype(np.float32)
    hard = np.abs(rng.standard_normal((F, T))).astype(np.float32)
    X = np.stack([easy, hard, easy * 0.5 + 0.01 * hard]).astype(np.float32)
    W, H, err, nit = ops.nmf_fit(dev(X), K, 120, 2e-3, 7, None, None)
    nits = [int(v) for v in nit]
    assert min(nits) < max(nits), nits                                  # at least one clip stopped before another
    import os
    os.environ["AINMF_NO_COOP"] = "1"           # one clip through the same batched kernels (not the cooperative one-clip launch)
    try:
        for b in range(3):
            Wb, Hb, eb, nb = ops.nmf_fit(dev(X[b:b + 1]), K, 120, 2e-3, 7, None, None)
            assert int(nb[0]) == nits[b]
            assert torch.equal(W[b], Wb[0]) and torch.equal(H[b], Hb[0]) and float(err[b]) == float(eb[0])
    finally:
        os.environ.pop("AINMF_NO_COOP")
    Wo, Ho, no, eo = libcalls.nmf_fit(X[1], K, seed=7, max_iter=120, tol=2e-3)
    assert abs(nits[1] - no) <= 2 and abs(float(err[1]) - eo) <= 1e-3 * eo
    # the same clip alone takes the cooperative launch (nmf_coop.cu: FFMA, stop rule on the device)
    Wc, Hc, ec, nc = ops.nmf_fit(dev(X[1:2]), K, 120, 2e-3, 7, None, None)
    assert abs(int(nc[0]) - no) <= 2 and abs(float(ec[0]) - eo) <= 1e-3 * eo


@pytest.mark.parametrize("F,T,K,iters,tol", [(1025, 863, 40, 30, 0.0), (513, 1724, 64, 25, 0.0), (257, 431, 100, 20, 0.0),
                                              (257, 500, 40, 300, 1e-3), (1025, 700, 128, 6, 0.0), (2049, 431, 64, 8, 0.0)])
def test_cooperative_one_clip_fit_matches_sklearn(ops, F, T, K, iters, tol):
    """One spectrogram = one cooperative launch (nmf_coop.cu): factors, objective and n_iter_ against sklearn's CD from the
    same seed, on the c2 / c3 shapes, K = 40 ... 128, an early stop, n_fft = 4096; and the same factors as the general kernels give (AINMF_NO_COOP=1) within rounding."""
    import os
    rng = np.random.default_rng(F + T + K)
    X = (np.abs(rng.standard_normal((F, 6))) @ np.abs(rng.standard_normal((6, T))) + 0.3 * np.abs(rng.standard_normal((F, T)))).astype(np.float32)
    X[:, 40:60] = X[:, :20].mean(axis=1, keepdims=True)                     # identical columns, as an imputed gap has
    Wo, Ho, no, eo = libcalls.nmf_fit(X, K, seed=3, max_iter=iters, tol=tol)
    launches0 = __import__("ainmf")._lib.lib().ainmf_launch_count()
    W, H, err, nit = ops.nmf_fit(dev(X[None]), K, iters, tol, 3, None, None)
    torch.cuda.synchronize()
    launches = __import__("ainmf")._lib.lib().ainmf_launch_count() - launches0
    assert launches < 40, launches                                          # not the per-iteration kernels
    assert abs(int(nit[0]) - no) <= (0 if tol == 0.0 else 1)
    assert abs(float(err[0]) - eo) <= 1e-4 * eo
    if tol == 0.0:
        assert rel_l2(W[0].cpu().numpy(), Wo) < 2e-3 and rel_l2(H[0].cpu().numpy(), Ho) < 2e-3
    os.environ["AINMF_NO_COOP"] = "1"
    try:
        W2, H2, err2, nit2 = ops.nmf_fit(dev(X[None]), K, iters, tol, 3, None, None)
    finally:
        os.environ.pop("AINMF_NO_COOP")
    if tol == 0.0:                                                          # with a stop rule the tf32 path may stop an iteration apart
        assert abs(float(err2[0]) - float(err[0])) <= 1e-4 * eo


def test_nmf_objective_is_monotone(ops):
    rng = np.random.default_rng(9)
    X = np.abs(rng.standard_normal((1, 513, 400))).astype(np.float32)
    errs = [float(ops.nmf_fit(dev(X), 64, it, 0.0, 1, None, None)[2][0]) for it in (1, 2, 4, 8, 16)]
    assert all(a >= b for a, b in zip(errs, errs[1:]))


# ---- end to end: golden vectors ----------------------------------------------------------------------------
def _run_columns(ops, x, **kw):
    p = dict(n_fft=1024, hop=256, rank=40, max_iter=200, tol=1e-4, seed=42, threshold=1e-4, frac_num=9, frac_den=10)
    p.update(kw)
    out = ops.nmf_inpaint(dev(x[None] if x.ndim == 1 else x), p["n_fft"], p["hop"], p["rank"], p["max_iter"], p["tol"],
                          p["seed"], p["threshold"], p["frac_num"], p["frac_den"], -1, -1, 1, None, None)
    return out


def test_c2_golden_end_to_end(ops, golden):
    """main4_NMF_gap.py on demo_assets/part2/damaged_gap.wav vs the shipped fixed_nmf_gap.wav and the oracle."""
    pcm = golden.gap_input_i16()
    x = ops.load_pcm16(dev(pcm[None]))[0]
    y, idx, nb, W, H, err, nit = ops.nmf_inpaint(x, 1024, 256, 40, 200, 1e-4, 42, 1e-4, 9, 10, -1, -1, 1, None, None)
    c2 = golden.c2
    n = int(nb[0])
    assert np.array_equal(idx[0, :n].cpu().numpy(), c2["bad_cols"])                 # bit-exact indices
    assert int(nit[0]) == int(c2["n_iter"]) == 200
    assert abs(float(err[0]) - float(c2["err"])) <= 1e-4 * float(c2["err"])         # objective within 1e-4
    yo, st = libcalls.restore_columns(x[0].cpu().numpy(), SR, return_all=True)
    yn = y[0].cpu().numpy()
    gs, ge = c2["gap"]
    assert libcalls.snr_db(yo, yn) >= 60.0                                         # north_star: >= 60 dB
    assert libcalls.snr_db(yo[gs:ge], yn[gs:ge]) >= 60.0
    shipped = pcm.astype(np.int32) + c2["shipped_minus_input_i16"]
    q = ops.store_pcm16(y[0]).cpu().numpy().astype(np.int32)
    outside = np.ones(len(q), bool)
    outside[gs - 1024:ge + 1024] = False
    assert np.max(np.abs(q[outside] - shipped[outside])) <= 1                       # untouched region: 1 LSB
    assert libcalls.snr_db(shipped.astype(np.float64), q.astype(np.float64)) >= 60.0


def test_c3_golden_end_to_end(ops, golden):
    pcm = golden.mask_input_i16()
    x = ops.load_pcm16(dev(pcm[None]))[0]
    y, idx, nb, W, H, err, nit = ops.nmf_inpaint(x, 1024, 256, 40, 200, 1e-4, 42, 0.01, 4, 5, -1, -1, 1, None, None)
    c3 = golden.c3
    assert np.array_equal(idx[0, :int(nb[0])].cpu().numpy(), c3["bad_cols"])
    assert int(nit[0]) == int(c3["n_iter"])
    assert abs(float(err[0]) - float(c3["err"])) <= 1e-4 * float(c3["err"])
    ref = pcm.astype(np.int32) + c3["ref_minus_input_i16"]
    q = ops.store_pcm16(y[0]).cpu().numpy().astype(np.int32)
    assert libcalls.snr_db(ref.astype(np.float64), q.astype(np.float64)) >= 60.0


def test_c2_at_2048_512_vs_oracle(ops, golden):
    """BASELINE config 2 names n_fft=2048 hop=512 (the script hard-codes 1024/256): checked against the oracle."""
    x = libcalls.load_normalised(golden.gap_input_i16())
    yo, st = libcalls.restore_columns(x, SR, n_fft=2048, hop=512, return_all=True)
    y, idx, nb, W, H, err, nit = _run_columns(ops, x, n_fft=2048, hop=512)
    assert np.array_equal(idx[0, :int(nb[0])].cpu().numpy(), st["bad"]) and int(nb[0]) == 172
    assert abs(float(err[0]) - st["err"]) <= 1e-4 * st["err"]
    gs, ge = golden.c2["gap"]
    yn = y[0].cpu().numpy()
    assert libcalls.snr_db(yo, yn) >= 60.0 and libcalls.snr_db(yo[gs:ge], yn[gs:ge]) >= 60.0


def test_c1_part0_golden(ops, golden):
    """main4_NMF.py: 50 refits with early stop, vs the shipped nmf_restored.wav and the reference run."""
    import ainmf
    c1 = golden.c1
    lab = ainmf.SpectralInpainter.__new__(ainmf.SpectralInpainter)
    ainmf.SpectralInpainter.__init__(lab, filename=None, duration=0.05)
    lab.sr = int(c1["sr"])
    lab.raw_audio = c1["raw"].copy()
    gs, ge = lab.apply_mask(0.2)
    assert (gs, ge) == tuple(c1["gap"]) and np.array_equal(lab.corrupted_audio, c1["corrupted"])
    out = lab.restore_with_nmf(n_components=40, n_iter=50)
    assert lab.cols_ == tuple(c1["cols"]) == (6, 10)
    assert libcalls.snr_db(c1["restored"], out) >= 60.0
    assert libcalls.snr_db(c1["restored"][gs:ge], out[gs:ge]) >= 60.0
    q = libcalls.quantise_int16(out).astype(np.int32)
    assert np.max(np.abs(q - c1["shipped_restored_i16"].astype(np.int32))) <= 1       # the oracle itself is within 1 LSB
    # the stop rule adds the violations in sklearn's order and precision where the decision is close (stop_kernel), so the
    # 50 chained refits stop where the reference's did
    assert lab.n_iter_ == int(c1["n_iters"][-1])
    # the final residual is 1e-4 of ||X||: its float32 cancellation noise, eps.||X|| / err = 4e-4, is in the reference's own
    # figure too, so the 1e-4 gate of the other configs (residual ~ 0.1 ||X||) is below what this number carries
    assert abs(lab.reconstruction_err_ - float(c1["err"])) <= 5e-4 * float(c1["err"])


# ---- edge cases --------------------------------------------------------------------------------------------
def test_good_first_frame_order_skips_bad_frames_exactly(ops):
    """With enough tiles to fill the machine the fit runs on a good-first permutation of the frames and skips the bad
    frames' share of both contractions (they are all the fill spectrum).  Same objective and waveform as without
    it and as the oracle; clips with many, few and no bad frames in one batch.  (The factors themselves are not compared:
    on this low-rank data H differs by 1e-4..1e-3 between ANY two summation orders, e.g. FFMA vs tensor path.)"""
    import os
    rng = np.random.default_rng(17)
    B, N, sr = 32, 66150, 44100
    t = np.arange(N) / sr
    X = np.empty((B, N), np.float32)
    for b in range(B):
        f = rng.uniform(200, 6000, 5)
        x = sum(np.sin(2 * np.pi * fi * t + rng.uniform(0, 6)) for fi in f) + 0.05 * rng.standard_normal(N)
        x = (x / np.abs(x).max()).astype(np.float32)
        if b % 3 == 0:
            x[20000:20000 + 4000 * (1 + b % 5)] = 0            # one gap of 0.09 .. 0.45 s
        elif b % 3 == 1:
            for s0 in rng.integers(0, N - 3000, 6):
                x[s0:s0 + int(rng.integers(600, 2500))] = 0     # scattered fragments
        X[b] = x                                                # b % 3 == 2: nothing bad
    args = (512, 128, 64, 30, 1e-4, 42, 1e-4, 9, 10, -1, -1, 1, None, None)
    os.environ.pop("AINMF_NO_COMPACT", None)
    y1, idx1, nb1, W1, H1, e1, n1 = ops.nmf_inpaint(dev(X), *args)
    os.environ["AINMF_NO_COMPACT"] = "1"
    try:
        y0, idx0, nb0, W0, H0, e0, n0 = ops.nmf_inpaint(dev(X), *args)
    finally:
        os.environ.pop("AINMF_NO_COMPACT", None)
    assert torch.equal(nb1, nb0) and torch.equal(idx1, idx0) and torch.equal(n1, n0)
    assert int((nb1 > 0).sum()) >= 20 and int((nb1 == 0).sum()) >= 8
    for b in range(B):
        if int(nb1[b]) == 0:
            assert torch.equal(y1[b], dev(X[b]))
            continue
        assert abs(float(e1[b]) - float(e0[b])) <= 1e-4 * float(e0[b])            # north_star tolerance on the objective
        assert libcalls.snr_db(y0[b].cpu().numpy(), y1[b].cpu().numpy()) > 90.0
    for b in (0, 1, 4):
        yo, st = libcalls.restore_columns(X[b], sr, n_fft=512, hop=128, K=64, seed=42, max_iter=30, return_all=True)
        assert int(nb1[b]) == len(st["bad"])
        assert abs(float(e1[b]) - st["err"]) <= 1e-4 * st["err"]
        assert libcalls.snr_db(yo, y1[b].cpu().numpy()) >= 60.0
    # the sum of the bad frames' rows of Ht runs on a side stream of the handle next to the X.Ht kernel (forked and joined
    # with events every iteration): same kernels, same order of additions -> bit-identical to the one-stream schedule, run
    # after run (a lost dependency between the two streams would show up here as a difference)
    os.environ["AINMF_NO_AUX_STREAM"] = "1"
    try:
        y2, idx2, nb2, W2, H2, e2, n2 = ops.nmf_inpaint(dev(X), *args)
    finally:
        os.environ.pop("AINMF_NO_AUX_STREAM", None)
    for _ in range(3):
        y3, idx3, nb3, W3, H3, e3, n3 = ops.nmf_inpaint(dev(X), *args)
        assert torch.equal(y3, y2) and torch.equal(W3, W2) and torch.equal(H3, H2) and torch.equal(e3, e2) and torch.equal(n3, n2)


def test_no_bad_frames_returns_input(ops):
    rng = np.random.default_rng(1)
    x = (0.5 + 0.1 * rng.standard_normal((2, 30000))).astype(np.float32)
    x[1, 10000:14000] = 0
    y, idx, nb, *_ = _run_columns(ops, x, max_iter=5)
    assert int(nb[0]) == 0 and int(nb[1]) > 0
    assert np.array_equal(y[0].cpu().numpy(), x[0])            # reference returns self.signal (main4_NMF_gap.py:54)
    assert not np.array_equal(y[1].cpu().numpy(), x[1])


def test_istft_blocks_without_modified_frames_return_the_input(ops):
    """Overlap-add identity used by the inverse: where no modified frame reaches, the windowed overlap-add of
    the unmodified frames is x * sum(w^2) and scipy divides by sum(w^2) (_spectral_py.py:1892-1910), so those samples
    equal the input; everything else still matches the oracle's full inverse transform."""
    sr, n, n_fft, hop = 16000, 160000, 1024, 256
    rng = np.random.default_rng(5)
    t = np.arange(n) / sr
    x = (0.4 * np.sin(2 * np.pi * 330.0 * t) + 0.2 * np.sin(2 * np.pi * 1234.0 * t) + 0.02 * rng.standard_normal(n)).astype(np.float32)
    x[70000:82000] = 0
    x /= np.abs(x).max()
    y, idx, nb, *_ = _run_columns(ops, x, rank=64, max_iter=20)
    y = y[0].cpu().numpy()
    bad = idx[0, :int(nb[0])].cpu().numpy()
    lo, hi = int(bad.min()) * hop - n_fft // 2, int(bad.max()) * hop + n_fft // 2      # samples the bad frames cover
    tile = 16 * hop                                                                     # kIstftHopsPerBlock * hop
    far_lo, far_hi = (lo // tile) * tile - tile, (hi // tile + 2) * tile
    assert far_lo > 0 and far_hi < n
    assert np.array_equal(y[:far_lo], x[:far_lo]) and np.array_equal(y[far_hi:], x[far_hi:])
    assert not np.array_equal(y[lo:hi], x[lo:hi])
    yo = libcalls.restore_columns(x, sr, n_fft=n_fft, hop=hop, K=64, seed=42, max_iter=20)
    assert libcalls.snr_db(yo, y) >= 60.0
    assert libcalls.snr_db(yo[:far_lo], y[:far_lo]) >= 100.0        # round trip of the oracle vs the exact identity


def test_invalid_arguments_raise(ops):
    import ainmf
    x = torch.zeros((1, 30000), device="cuda")
    with pytest.raises(ainmf.AinmfError):
        ops.stft(x, 1000, 250)                                  # not a power of two
    with pytest.raises(ainmf.AinmfError):
        ops.stft(torch.zeros((1, 100), device="cuda"), 1024, 256)   # shorter than one frame
    with pytest.raises(ainmf.AinmfError):
        _run_columns(ops, x.cpu().numpy(), rank=500)
    with pytest.raises(ainmf.AinmfError) as e:
        _run_columns(ops, np.zeros(30720, np.float32))          # every frame silent (30720 = 120*hop) -> fill undefined
    assert e.value.code == -5
    with pytest.raises(ainmf.AinmfError) as e:
        ops.stft(torch.zeros((1, 20000), device="cuda"), 4096, 4096)   # > 227 KB of shared memory per block: rejected up front
    assert "shared memory" in str(e.value)
    with pytest.raises(NotImplementedError):
        ops.stft(torch.zeros((1, 4096)), 1024, 256)             # CPU tensor: no CPU implementation


def test_all_silent_clip_inside_a_batch_is_passed_through(ops):
    """A clip whose every frame is flagged has no fill spectrum (NaN in the reference).  Alone it fails the call
    (AINMF_ERR_ALL_BAD); inside a batch it is returned unchanged with err = NaN, n_iter = 0, n_bad = T, and the other
    clips are restored exactly as they are without it."""
    rng = np.random.default_rng(3)
    N = 30720
    x = (0.4 * np.sin(2 * np.pi * 440.0 * np.arange(N) / 16000.0) + 0.05 * rng.standard_normal(N)).astype(np.float32)
    x[12000:16000] = 0
    X = np.stack([x, np.zeros(N, np.float32), np.roll(x, 5000)])
    y, idx, nb, W, H, err, nit = _run_columns(ops, X, max_iter=12)
    T = idx.shape[1]
    assert int(nb[1]) == T and int(nit[1]) == 0 and np.isnan(float(err[1]))
    assert np.array_equal(y[1].cpu().numpy(), X[1])
    assert np.array_equal(idx[1].cpu().numpy(), np.arange(T))
    for b in (0, 2):
        yb, ib, nbb, _, _, eb, nib = _run_columns(ops, X[b], max_iter=12)
        assert torch.equal(y[b], yb[0]) and float(err[b]) == float(eb[0]) and int(nit[b]) == int(nib[0]) == 12
        n = int(nb[b])
        assert n == int(nbb[0]) and torch.equal(idx[b, :n], ib[0, :n]) and bool((idx[b, n:] == -1).all())


def test_host_entry_point_pipelines_chunks_and_matches_the_device_op(ops):
    """ainmf_inpaint_host splits a batch into chunks with two in flight (copy-in | fit | copy-out on three streams).  Clips
    are independent, so the chunked host call must return for every clip what one device call on the whole batch returns --
    same frame counts and iteration counts, objective within 1e-5, waveform within 90 dB (the kernels pick their split
    factors from the batch size, so the floating-point summation order differs between a 300-clip and a 150-clip launch;
    a lost dependency between the three streams would show up as garbage, not as 1e-6) -- for the default split
    (300 clips -> 148 + 152) and for a memory cap that forces many small chunks with buffer reuse."""
    import ctypes as C
    import ainmf
    from ainmf import _capi
    rng = np.random.default_rng(8)
    B, N = 300, 12000
    t = np.arange(N) / 8000.0
    X = np.empty((B, N), np.float32)
    for b in range(B):
        x = np.sin(2 * np.pi * rng.uniform(200, 1500) * t) + 0.3 * np.sin(2 * np.pi * rng.uniform(1500, 3500) * t) + 0.05 * rng.standard_normal(N)
        x = (x / np.abs(x).max()).astype(np.float32)
        if b % 7 != 3:                                   # every seventh clip has nothing to restore
            s0 = int(rng.integers(500, N - 3000))
            x[s0:s0 + int(rng.integers(300, 2000))] = 0
        X[b] = x
    yd, idx, nbd, W, H, errd, nitd = ops.nmf_inpaint(dev(X), 256, 64, 16, 25, 0.0, 42, 1e-4, 9, 10, -1, -1, 1, None, None)
    L = ainmf._lib.lib()
    h = ainmf._lib.handle(0)
    p = _capi.default_params(L, batch=B, n_samples=N, n_fft=256, hop=64, rank=16, max_iter=25, tol=0.0, seed=42,
                             threshold=1e-4, frac_num=9, frac_den=10)
    xh = torch.from_numpy(X).pin_memory()
    ydn, errn = yd.cpu().numpy(), errd.cpu().numpy()
    for cap in (0, 40 << 20):                            # default chunking (148 + 152); 40 MB cap -> many chunks, buffers reused
        yh = torch.zeros((B, N), dtype=torch.float32).pin_memory()
        nb, er, ni = np.zeros(B, np.int32), np.zeros(B, np.float32), np.zeros(B, np.int32)
        rc = L.ainmf_inpaint_host(h, C.byref(p), C.c_void_p(xh.data_ptr()), C.c_void_p(yh.data_ptr()), nb.ctypes.data_as(C.c_void_p),
                                  er.ctypes.data_as(C.c_void_p), ni.ctypes.data_as(C.c_void_p), cap)
        ainmf._lib.check(rc, 0)
        assert np.array_equal(nb, nbd.cpu().numpy()) and np.array_equal(ni, nitd.cpu().numpy())
        yn = yh.numpy()
        for b in range(B):
            if nb[b] == 0:
                assert np.array_equal(yn[b], X[b])
            else:
                assert abs(er[b] - errn[b]) <= 1e-5 * errn[b]
                assert libcalls.snr_db(ydn[b], yn[b]) >= 90.0
    assert int((nbd == 0).sum()) >= 30 and int((nbd > 0).sum()) >= 200


def test_host_entry_point_pcm16_is_the_file_to_file_path(ops, golden):
    """ainmf_inpaint_host_pcm16: int16 samples as wavfile.read returns them -> int16 samples as wavfile.write receives them
    (load_damaged_data -> restore -> save_result on the device, inside the chunk pipeline).  (1) The shipped damaged_gap.wav
    gives the shipped fixed_nmf_gap.wav within 1 LSB outside the gap and 60 dB overall, the peak and the bad columns of the
    reference.  (2) A 300-clip stereo batch (two chunks, some clips with nothing to restore) equals, sample for sample up to
    1 LSB on at most 0.1 % of the samples, what the device ops give on the whole batch (load_pcm16 -> nmf_inpaint ->
    store_pcm16; the chunked call uses other split factors, so the float sums differ in the last bits), and exactly for the
    clips that have no bad frame."""
    import ctypes as C
    import ainmf
    from ainmf import _capi
    L = ainmf._lib.lib()
    h = ainmf._lib.handle(0)
    vp = C.c_void_p
    # (1) the golden file
    pcm = golden.gap_input_i16()
    N = len(pcm)
    pin = torch.from_numpy(pcm.copy()).pin_memory()
    pout = torch.zeros(N, dtype=torch.int16).pin_memory()
    peak, nb, er, ni = np.zeros(1, np.float32), np.zeros(1, np.int32), np.zeros(1, np.float32), np.zeros(1, np.int32)
    p = _capi.default_params(L, batch=1, n_samples=N, n_fft=1024, hop=256, rank=40, max_iter=200, tol=1e-4, seed=42,
                             threshold=1e-4, frac_num=9, frac_den=10)
    rc = L.ainmf_inpaint_host_pcm16(h, C.byref(p), vp(pin.data_ptr()), 1, vp(pout.data_ptr()), peak.ctypes.data_as(vp),
                                    nb.ctypes.data_as(vp), er.ctypes.data_as(vp), ni.ctypes.data_as(vp), 0)
    ainmf._lib.check(rc, 0)
    c2 = golden.c2
    assert float(peak[0]) == float(np.max(np.abs(pcm))) and int(nb[0]) == len(c2["bad_cols"]) and int(ni[0]) == 200
    assert abs(float(er[0]) - float(c2["err"])) <= 1e-4 * float(c2["err"])
    q = pout.numpy().astype(np.int32)
    shipped = pcm.astype(np.int32) + c2["shipped_minus_input_i16"]
    gs, ge = c2["gap"]
    outside = np.ones(N, bool)
    outside[gs - 1024:ge + 1024] = False
    assert np.max(np.abs(q[outside] - shipped[outside])) <= 1
    assert libcalls.snr_db(shipped.astype(np.float64), q.astype(np.float64)) >= 60.0
    # (2) a stereo batch in two chunks against the device ops
    rng = np.random.default_rng(11)
    B, N = 300, 12000
    t = np.arange(N) / 8000.0
    P = np.empty((B, N, 2), np.int16)
    for b in range(B):
        x = np.sin(2 * np.pi * rng.uniform(200, 1500) * t) + 0.3 * np.sin(2 * np.pi * rng.uniform(1500, 3500) * t) + 0.05 * rng.standard_normal(N)
        x = x / np.abs(x).max() * rng.uniform(0.3, 0.95)
        if b % 7 != 3:
            s0 = int(rng.integers(500, N - 3000))
            x[s0:s0 + int(rng.integers(300, 2000))] = 0
        P[b, :, 0] = np.round(x * 32767).astype(np.int16)
        P[b, :, 1] = np.round(0.8 * x * 32767).astype(np.int16)
    xd, pkd = ops.load_pcm16(dev(P))
    yd, idx, nbd, W, H, errd, nitd = ops.nmf_inpaint(xd, 256, 64, 16, 25, 0.0, 42, 1e-4, 9, 10, -1, -1, 1, None, None)
    qd = ops.store_pcm16(yd).cpu().numpy()
    pin = torch.from_numpy(P).pin_memory()
    pout = torch.zeros((B, N), dtype=torch.int16).pin_memory()
    peak, nb, er, ni = np.zeros(B, np.float32), np.zeros(B, np.int32), np.zeros(B, np.float32), np.zeros(B, np.int32)
    p = _capi.default_params(L, batch=B, n_samples=N, n_fft=256, hop=64, rank=16, max_iter=25, tol=0.0, seed=42,
                             threshold=1e-4, frac_num=9, frac_den=10)
    for cap in (0, 40 << 20):
        pout.zero_()
        rc = L.ainmf_inpaint_host_pcm16(h, C.byref(p), vp(pin.data_ptr()), 2, vp(pout.data_ptr()), peak.ctypes.data_as(vp),
                                        nb.ctypes.data_as(vp), er.ctypes.data_as(vp), ni.ctypes.data_as(vp), cap)
        ainmf._lib.check(rc, 0)
        assert np.array_equal(peak, pkd.cpu().numpy()) and np.array_equal(nb, nbd.cpu().numpy()) and np.array_equal(ni, nitd.cpu().numpy())
        q = pout.numpy()
        d = np.abs(q.astype(np.int32) - qd.astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() <= 1e-3
        clean = nb == 0
        assert clean.sum() >= 30 and np.array_equal(q[clean], qd[clean])


def test_shim_classes_match_reference_interface(ops, golden, tmp_path):
    import ainmf
    from scipy.io import wavfile
    path = tmp_path / "damaged_gap.wav"
    wavfile.write(path, SR, golden.gap_input_i16())
    lab = ainmf.NMFFairGapInpainter(str(path), output_dir=str(tmp_path))
    assert lab.restore() is None                                # not loaded yet (main4_NMF_gap.py:43)
    lab.load_damaged_data()
    assert lab.sr == SR and lab.signal.dtype == np.float32 and len(lab.signal) == 441000
    bad = lab.get_gap_mask(1724, 256)
    assert bad.dtype == np.int64 and np.array_equal(bad, golden.c2["bad_cols"])
    res = lab.restore()
    assert res.dtype == np.float32 and res.shape == (441000,)
    lab.save_result(res)
    _, saved = wavfile.read(tmp_path / "fixed_nmf_gap.wav")
    shipped = golden.gap_input_i16().astype(np.int32) + golden.c2["shipped_minus_input_i16"]
    assert libcalls.snr_db(shipped.astype(np.float64), saved.astype(np.float64)) >= 60.0
    m = ainmf.NMFFairInpainter(str(path))
    m.load_damaged_data()
    assert np.array_equal(m.get_mask_from_signal(1724, 256), libcalls.column_mask(m.signal, 1724, 256, 0.01, 0.8))


@pytest.mark.parametrize("F,T,K,iters,tol", [(513, 300, 64, 30, 0.0), (257, 400, 40, 60, 1e-3), (129, 200, 16, 20, 0.0)])
def test_mu_solver_matches_sklearn_mu(ops, F, T, K, iters, tol):
    """solver='mu' (Frobenius multiplicative update, the north-star-shaped solver) vs sklearn solver='mu'."""
    rng = np.random.default_rng(F + K)
    X = np.abs(rng.standard_normal((F, T))).astype(np.float32)
    W0, Ht0 = restate.init_factors(X.mean(), F, T, K, 3)
    Wo, Ho, no, eo = libcalls.nmf_fit(X, K, W0=W0, H0=Ht0.T, max_iter=iters, tol=tol, solver="mu")
    W, H, err, nit = ops.nmf_fit(dev(X[None]), K, iters, tol, 0, dev(W0[None]), dev(np.ascontiguousarray(Ht0.T)[None]), "mu")
    assert int(nit[0]) == no
    assert abs(float(err[0]) - eo) <= 1e-4 * eo
    assert rel_l2(W[0].cpu().numpy(), Wo) < 1e-3 and rel_l2(H[0].cpu().numpy(), Ho) < 1e-3


@pytest.mark.parametrize("F,T,K,iters,tol", [(129, 200, 16, 25, 0.0), (513, 431, 40, 12, 0.0), (1025, 300, 128, 6, 0.0),
                                              (100, 257, 24, 200, 1e-3)])
def test_mu_kl_solver_matches_sklearn(ops, F, T, K, iters, tol):
    """solver='mu-kl' -- north_star (3)'s form: W.H, the ratio X / (W.H) and both contractions in one fused kernel per
    half-step (nmf_mukl.cu) -- against sklearn solver='mu', beta_loss='kullback-leibler' from the same initial factors:
    n_iter_ (every-10th-iteration test), reconstruction_err_ = sqrt(2 D_KL) within 1e-4, factors within 1e-3.  Zero rows
    and columns exercise the x <= eps and W.H < eps branches."""
    rng = np.random.default_rng(F + K)
    X = np.abs(rng.standard_normal((F, T))).astype(np.float32)
    X[:, 7:11] = 0.0
    X[3, :] = 0.0
    W0, Ht0 = restate.init_factors(X.mean(), F, T, K, 3)
    Wo, Ho, no, eo = libcalls.nmf_fit(X, K, W0=W0, H0=Ht0.T, max_iter=iters, tol=tol, solver="mu", beta_loss="kullback-leibler")
    W, H, err, nit = ops.nmf_fit(dev(X[None]), K, iters, tol, 0, dev(W0[None]), dev(np.ascontiguousarray(Ht0.T)[None]), "mu-kl")
    assert int(nit[0]) == no
    assert abs(float(err[0]) - eo) <= 1e-4 * eo
    assert rel_l2(W[0].cpu().numpy(), Wo) < 1e-3 and rel_l2(H[0].cpu().numpy(), Ho) < 1e-3


def test_mu_kl_batch_is_independent(ops):
    """A batch of clips through the fused KL kernels equals the clips fitted one by one, bit for bit."""
    rng = np.random.default_rng(5)
    X = np.abs(rng.standard_normal((3, 257, 150))).astype(np.float32)
    Wb, Hb, eb, nb = ops.nmf_fit(dev(X), 20, 15, 0.0, 7, None, None, "mu-kl")
    for b in range(3):
        W1, H1, e1, n1 = ops.nmf_fit(dev(X[b:b + 1]), 20, 15, 0.0, 7, None, None, "mu-kl")
        assert torch.equal(W1[0], Wb[b]) and torch.equal(H1[0], Hb[b]) and float(e1[0]) == float(eb[b])


def test_ffma_path_still_matches_when_tensor_cores_disabled(ops, golden):
    """AINMF_DISABLE_TC=1 routes K >= 64 problems through the FFMA kernels; both paths must satisfy the same gates."""
    import os
    import subprocess
    import sys
    code = ("import numpy as np, torch, sys; sys.path.insert(0, %r); import ainmf; from oracle import libcalls; "
            "rng = np.random.default_rng(1); X = np.abs(rng.standard_normal((513, 300))).astype(np.float32); "
            "Wo, Ho, no, eo = libcalls.nmf_fit(X, 64, seed=0, max_iter=15, tol=0.0); "
            "W, H, err, nit = ainmf.ops.nmf_fit(torch.from_numpy(X[None]).cuda(), 64, 15, 0.0, 0, None, None); "
            "assert abs(float(err[0]) - eo) <= 1e-4 * eo, (float(err[0]), eo); print('FFMA_OK')") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, AINMF_DISABLE_TC="1"), capture_output=True, text=True, timeout=300)
    assert "FFMA_OK" in out.stdout, out.stdout[-500:] + out.stderr[-1500:]
