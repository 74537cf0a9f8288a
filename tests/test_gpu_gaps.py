"""Callers / baselines either side of the NMF path: sample-level gap detectors, linear interpolation,
Part-0 blend and SNR -- GPU kernels vs the numpy restatement of the sibling scripts (oracle/libcalls.py)."""
import numpy as np
import pytest

from oracle import libcalls

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def signals():
    rng = np.random.default_rng(7)
    out = []
    for N, kind in ((5000, "gaps"), (44100, "gaps"), (100003, "edges"), (3000, "none"), (2049, "all")):
        x = (rng.standard_normal(N) * 0.3).astype(np.float32)
        if kind == "gaps":
            for _ in range(12):
                s = int(rng.integers(0, N - 500)); l = int(rng.integers(1, 450))
                x[s:s + l] = 0
            x[rng.integers(0, N, 50)] = 5e-5                    # below 1e-4: isolated damaged samples
        elif kind == "edges":
            x[:300] = 0; x[-1234:] = 0; x[5000:5101] = 0; x[7000:7100] = 0     # run of exactly 101 and of 100 samples
            x[40960:43008] = 0                                  # a run covering whole 2048-sample chunks
        elif kind == "all":
            x[:] = 0
        out.append(x)
    return out


@pytest.mark.parametrize("thr", [1e-4, 0.01])
def test_find_main_gap_and_find_gaps_bit_exact(thr):
    import ainmf
    for x in signals():
        span = ainmf.ops.find_main_gap(dev(x[None]), thr)[0].cpu().numpy()
        ref = libcalls.find_main_gap(x, thr)
        assert (tuple(span) == ref) if ref is not None else (tuple(span) == (-1, -1))
        runs, n = ainmf.ops.find_gaps(dev(x[None]), thr, 100, 64)
        ref_runs = libcalls.find_gaps(x, thr, 100)
        assert int(n[0]) == len(ref_runs)
        assert [tuple(r) for r in runs[0, :len(ref_runs)].cpu().numpy()] == ref_runs


def test_find_gaps_batch_and_overflow():
    import ainmf
    xs = [x for x in signals() if len(x) == 5000] * 3
    X = np.stack(xs)
    X[1, 100:400] = 0.3
    runs, n = ainmf.ops.find_gaps(dev(X), 0.01, 100, 2)          # fewer slots than runs: count is still the true count
    for b in range(3):
        ref = libcalls.find_gaps(X[b], 0.01, 100)
        assert int(n[b]) == len(ref)
        assert [tuple(r) for r in runs[b, :min(2, len(ref))].cpu().numpy()] == ref[:2]


def test_linear_interp_matches_numpy():
    import ainmf
    for x in signals():
        y, nd = ainmf.ops.linear_interp(dev(x[None]), 1e-4)
        yo, no = libcalls.linear_interp(x, 1e-4)
        assert int(nd[0]) == no
        y = y[0].cpu().numpy()
        valid = np.abs(x) > 1e-4
        assert np.array_equal(y[valid], x[valid])
        # np.interp runs in float64 and the result is rounded to float32: same operations here -> at most 1 ulp apart
        assert np.max(np.abs(y.astype(np.float64) - yo.astype(np.float64))) <= 1.2e-7 * max(1.0, float(np.max(np.abs(yo))))


def test_blend_and_snr_match_part0(golden):
    import ainmf
    rng = np.random.default_rng(3)
    raw = (rng.standard_normal(2205) * 0.4).astype(np.float32)
    res = (raw + 0.05 * rng.standard_normal(2205)).astype(np.float32)
    gs, ge = 882, 1323
    out = ainmf.ops.blend_boundaries(dev(raw), dev(res), gs, ge, 50).cpu().numpy()
    ref = libcalls.part0_blend(raw, res, gs, ge)
    assert np.array_equal(out, ref)                             # float64 ramp arithmetic reproduced exactly
    for (b, e) in ((0, 2205), (gs, ge)):
        s = ainmf.ops.snr_db(dev(raw), dev(out), b, e)
        assert abs(s - libcalls.snr_db(raw[b:e], ref[b:e])) < 1e-3


def test_apply_gaps_matches_generate_part1():
    import ainmf
    rng = np.random.default_rng(5)
    N = 441000
    x = (rng.standard_normal((2, N)) * 0.3 + 1.0).astype(np.float32)
    starts = np.empty((2, 551), np.int64); lens = np.empty((2, 551), np.int64)
    ref = x.copy()
    for b in range(2):
        np.random.seed(b)
        for g in range(551):                        # generate_part1_data.create_random_mask(N, 0.25, 400)
            l = np.random.randint(50, 400); s = np.random.randint(0, N - l)
            ref[b, s:s + l] = 0
            starts[b, g], lens[b, g] = s, l
    out = ainmf.ops.apply_gaps_(dev(x), dev(starts), dev(lens)).cpu().numpy()
    assert np.array_equal(out, ref)


def test_f4_golden_outputs_of_the_reference_scripts(golden):
    """Device kernels vs what the unmodified sibling scripts produced (tests/golden/f4_siblings.npz)."""
    import ainmf
    f4 = golden.f4
    x_gap = libcalls.load_normalised(golden.gap_input_i16())
    span = ainmf.ops.find_main_gap(dev(x_gap[None]), 1e-4)[0].cpu().numpy()
    assert tuple(span) == tuple(f4["main_gap"])
    dr = f4["damaged_random_i16"]
    runs, n = ainmf.ops.find_gaps(dev(libcalls.load_normalised(dr)[None]), 0.01, 100, 256)
    assert int(n[0]) == len(f4["gaps_random"]) and np.array_equal(runs[0, :int(n[0])].cpu().numpy(), f4["gaps_random"])
    xr = dr.astype(np.float32) / np.max(np.abs(dr))
    y, nd = ainmf.ops.linear_interp(dev(xr[None]), 1e-4)
    assert int(nd[0]) == int(f4["n_damaged"])
    pcm = ainmf.ops.store_pcm16(y[0]).cpu().numpy()
    assert np.max(np.abs(pcm.astype(np.int32) - f4["fixed_linear_i16"].astype(np.int32))) <= 1      # int16 LSB (1-ulp float ties)
