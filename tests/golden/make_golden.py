"""Generate tests/golden/*.npz by running the UNMODIFIED reference scripts (build container only).

    python tests/golden/make_golden.py

Needs /root/reference (read-only) -- it is not available on the GPU box, which is why the
outputs are committed.  Every fixture records both what the reference script produced here and,
where the reference ships one, the shipped WAV it must agree with to <= 1 int16 LSB.
"""
from __future__ import annotations

import os
import sys

import numpy as np
from scipy.io import wavfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import libcalls, ref_loader  # noqa: E402

REF = ref_loader.REF
OUT = os.path.dirname(os.path.abspath(__file__))


def rd(rel):
    return wavfile.read(os.path.join(REF, rel))


def main():
    assert ref_loader.available(), "reference tree not found"
    sr, base = rd("vocals_accompaniment_10s.wav")
    _, original = rd("demo_assets/part2/original.wav")
    _, damaged_gap = rd("demo_assets/part2/damaged_gap.wav")
    _, fixed_gap = rd("demo_assets/part2/fixed_nmf_gap.wav")

    # ---- the clean 10 s clip, as generate_part2_data.py:29-34,58-60 makes it ----------------
    data = base.mean(axis=1)
    data = data.astype(np.float32) / np.max(np.abs(data))
    data = data[:10 * sr]
    assert np.array_equal(libcalls.quantise_int16(data), original)
    gs, ge = libcalls.centre_gap(len(data), sr)
    cor = data.copy()
    cor[gs:ge] = 0
    assert np.array_equal(libcalls.quantise_int16(cor), damaged_gap), "damaged_gap.wav not reproduced"
    np.savez_compressed(os.path.join(OUT, "clip10s.npz"), original_i16=original, sr=sr)

    # ---- config 2: main4_NMF_gap.py, unmodified, on the shipped damaged_gap.wav --------------
    with ref_loader.scratch_cwd({"demo_assets/part2/damaged_gap.wav":
                                 os.path.join(REF, "demo_assets/part2/damaged_gap.wav")}):
        g = ref_loader.load("main4_NMF_gap")
        res = np.asarray(g.res)
        x = g.lab.signal
        bad = g.lab.get_gap_mask(1724, 256)
        regenerated = wavfile.read("demo_assets/part2/fixed_nmf_gap.wav")[1]
    y, st = libcalls.restore_columns(x, sr, return_all=True)
    assert np.array_equal(y, res), "oracle.libcalls differs from the reference script"
    lsb = int(np.max(np.abs(regenerated.astype(np.int32) - fixed_gap.astype(np.int32))))
    print("C2: regenerated vs shipped fixed_nmf_gap.wav max |diff| =", lsb, "LSB; n_bad", len(bad),
          "n_iter", st["n_iter"], "err", st["err"])
    assert lsb <= 1
    np.savez_compressed(
        os.path.join(OUT, "c2_gap.npz"),
        bad_cols=bad.astype(np.int64), n_iter=st["n_iter"], err=np.float64(st["err"]),
        shipped_minus_input_i16=(fixed_gap.astype(np.int32) - damaged_gap.astype(np.int32)).astype(np.int32),
        ref_minus_shipped_i16=(regenerated.astype(np.int32) - fixed_gap.astype(np.int32)).astype(np.int8),
        mag_sub=st["mag"][::16, ::16].astype(np.float32),       # (F,T) order, every 16th bin/frame
        y_sub=res[::32].astype(np.float32),
        y_gap_f16=res[gs:ge].astype(np.float16),                 # coarse copy of the in-gap waveform
        W_fro=np.float64(np.linalg.norm(st["W"])), H_fro=np.float64(np.linalg.norm(st["H"])),
        gap=np.array([gs, ge]), shipped_lsb=lsb)

    # ---- config 3: main4_NMF_mask.py, unmodified, on a seeded generate_part1_data input ------
    np.random.seed(0)
    keep = libcalls.create_random_mask(len(data), mask_ratio=0.25)
    cor3 = data.copy()
    cor3[~keep] = 0
    cor3_i16 = libcalls.quantise_int16(cor3)
    assert np.array_equal(cor3_i16, np.where(keep, original, 0))
    with ref_loader.scratch_cwd() as d:
        os.makedirs("demo_assets", exist_ok=True)
        wavfile.write("demo_assets/damaged_random.wav", sr, cor3_i16)
        m = ref_loader.load("main4_NMF_mask")
        res3 = np.asarray(m.res)
        x3 = m.lab.signal
        bad3 = m.lab.get_mask_from_signal(1724, 256)
        out3 = wavfile.read("demo_assets/fixed_nmf_random.wav")[1]
    y3, st3 = libcalls.restore_columns(x3, sr, threshold=0.01, frac=0.8, return_all=True)
    assert np.array_equal(y3, res3)
    print("C3: n_bad", len(bad3), "n_iter", st3["n_iter"], "err", st3["err"])
    np.savez_compressed(
        os.path.join(OUT, "c3_mask.npz"),
        keep_bits=np.packbits(keep), bad_cols=bad3.astype(np.int64), n_iter=st3["n_iter"],
        err=np.float64(st3["err"]),
        ref_minus_input_i16=(out3.astype(np.int32) - cor3_i16.astype(np.int32)).astype(np.int32),
        y_sub=res3[::32].astype(np.float32), seed=0)

    # ---- config 1: main4_NMF.py __main__ body (:163-170) ------------------------------------
    with ref_loader.scratch_cwd({"vocals_accompaniment_10s.wav":
                                 os.path.join(REF, "vocals_accompaniment_10s.wav")}):
        p0 = ref_loader.load("main4_NMF")
        lab = p0.SpectralInpainter(filename="vocals_accompaniment_10s.wav", duration=0.05)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            lab.load_data()
            gs0, ge0 = lab.apply_mask(gap_ratio=0.2)
            lab.restore_with_nmf(n_components=40, n_iter=50)
    ship = {k: rd(f"demo_assets/part0/nmf_{k}.wav")[1] for k in ("original", "corrupted", "restored")}
    assert np.array_equal(libcalls.quantise_int16(lab.raw_audio), ship["original"])
    assert np.array_equal(libcalls.quantise_int16(lab.corrupted_audio), ship["corrupted"])
    lsb0 = int(np.max(np.abs(libcalls.quantise_int16(lab.restored_audio).astype(np.int32)
                             - ship["restored"].astype(np.int32))))
    raw = libcalls.part0_load(base, sr, 0.05)
    assert np.array_equal(raw, lab.raw_audio)
    cor0, a, b = libcalls.part0_apply_mask(raw, 0.2)
    assert (a, b) == (gs0, ge0) and np.array_equal(cor0, lab.corrupted_audio)
    y0, st0 = libcalls.part0_restore(raw, cor0, sr, a, b, return_all=True)
    assert np.array_equal(y0, lab.restored_audio), "oracle.libcalls part0 differs from the reference"
    print("C1: restored vs shipped max |diff| =", lsb0, "LSB; cols", st0["cols"], "n_iters", st0["n_iters"][:6],
          "...", st0["n_iters"][-1], "err", st0["err"])
    assert lsb0 <= 1
    np.savez_compressed(
        os.path.join(OUT, "c1_part0.npz"),
        raw=lab.raw_audio.astype(np.float32), corrupted=lab.corrupted_audio.astype(np.float32),
        restored=lab.restored_audio.astype(np.float32), pre_blend=st0["pre_blend"].astype(np.float32),
        shipped_original_i16=ship["original"], shipped_corrupted_i16=ship["corrupted"],
        shipped_restored_i16=ship["restored"], gap=np.array([gs0, ge0]), cols=np.array(st0["cols"]),
        n_iters=np.array(st0["n_iters"]), err=np.float64(st0["err"]), sr=sr, shipped_lsb=lsb0)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()


def siblings():
    """the sample-level detectors of main3_AR_text_gap.py / main3_AR_text_mask.py (class bodies executed
    unmodified; the module tails that run the AR restoration are not) and linear_interp_part1.py run as shipped."""
    import contextlib
    import io

    def class_namespace(script):
        src = open(os.path.join(REF, script)).read()
        head = src[:src.index("\nlab = ")]                   # everything before the module tail
        ref_loader._stub_matplotlib()
        ns = {}
        exec(compile(head, os.path.join(REF, script), "exec"), ns)
        return ns

    sr, dg = rd("demo_assets/part2/damaged_gap.wav")
    x_gap = libcalls.load_normalised(dg)
    _, dr = rd("demo_assets/part1/damaged_random.wav")
    x_rand = libcalls.load_normalised(dr)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        ns = class_namespace("main3_AR_text_gap.py")
        a = ns["ARFairGapInpainter"].__new__(ns["ARFairGapInpainter"])
        a.signal = x_gap
        span = a.find_main_gap()
        ns = class_namespace("main3_AR_text_mask.py")
        m = ns["IterativeARInpainter"].__new__(ns["IterativeARInpainter"])
        m.signal = x_rand
        gaps = m.find_gaps()
        m.signal = x_gap
        gaps_on_gap = m.find_gaps()
        with ref_loader.scratch_cwd({"demo_assets/part1/damaged_random.wav": os.path.join(REF, "demo_assets/part1/damaged_random.wav")}):
            li = ref_loader.load("linear_interp_part1")
            li.linear_interpolation_restoration()
            fixed = wavfile.read("demo_assets/part1/fixed_linear_random.wav")[1]
    assert tuple(span) == libcalls.find_main_gap(x_gap, 1e-4)
    assert [tuple(g) for g in gaps] == libcalls.find_gaps(x_rand, 0.01, 100)
    assert [tuple(g) for g in gaps_on_gap] == libcalls.find_gaps(x_gap, 0.01, 100)
    # linear_interp_part1.py normalises without the mono/zero guards of the loaders: data.astype(float32) / max|data|
    xr = dr.astype(np.float32) / np.max(np.abs(dr))
    y, nd = libcalls.linear_interp(xr, 1e-4)
    assert np.array_equal(libcalls.quantise_int16(y), fixed), "oracle linear_interp differs from linear_interp_part1.py"
    np.savez_compressed(os.path.join(OUT, "f4_siblings.npz"), main_gap=np.array(span, np.int64),
                        gaps_random=np.array(gaps, np.int64).reshape(-1, 2), gaps_on_gap=np.array(gaps_on_gap, np.int64).reshape(-1, 2),
                        damaged_random_i16=dr, fixed_linear_i16=fixed, n_damaged=np.int64(nd))
    print("F4: main gap", tuple(span), "| find_gaps on damaged_random:", len(gaps), "runs | linear interp:", nd, "damaged samples, int16 output bit-equal")


if __name__ == "__main__" and "--siblings" in sys.argv:
    siblings()
