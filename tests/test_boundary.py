"""The drop-in boundary without a GPU: the CUDA library builds, loads and exports every symbol the header
declares; host-only entry points agree with the oracle; the product path refuses to run on the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import restate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import ainmf
    from ainmf import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ainmf_build", os.path.join(ROOT, "audio-inpainting_b200", "build.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        m.build_library()
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ainmf.h")).read()
    names = sorted(set(re.findall(r"\b(ainmf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    import ainmf
    assert sorted(ainmf._capi.EXPORTS) == names
    for n in names:
        assert hasattr(lib, n), n


def test_host_chunk_schedule(lib):
    """Host logic of ainmf_inpaint_host / _pcm16: every clip in exactly one chunk, no chunk above the cap, a first chunk of
    one clip per SM when the batch holds two such chunks, multiples of the SM count in the middle, no short tail."""
    import ctypes as C
    import numpy as np

    def sched(batch, cap, sm=148):
        sizes = np.zeros(4096, np.int32)
        n = C.c_int32(0)
        assert lib.ainmf_host_chunk_schedule(batch, cap, sm, sizes.ctypes.data_as(C.c_void_p), len(sizes), C.byref(n)) == 0
        return sizes[:n.value].tolist()

    assert sched(1, 512) == [1] and sched(100, 512) == [100] and sched(295, 512) == [295]
    assert sched(296, 512) == [148, 148] and sched(300, 512) == [148, 152]
    assert sched(512, 512) == [148, 364]                                   # the bench's batch
    assert sched(4096, 512) == [148] + [444] * 8 + [396]
    assert sched(668, 512) == [148, 296, 224]                              # 444 would leave a tail of 76
    assert sched(300, 17) == [17] * 17 + [11]                              # memory cap below one clip per SM
    assert lib.ainmf_host_chunk_schedule(0, 512, 148, None, 0, None) != 0
    rng = np.random.default_rng(0)
    for _ in range(300):
        batch, cap, sm = int(rng.integers(1, 6000)), int(rng.integers(1, 700)), int(rng.choice([8, 132, 148, 160]))
        s = sched(batch, cap, sm)
        assert sum(s) == batch and min(s) >= 1 and max(s) <= min(cap, 512), (batch, cap, sm, s)
        if cap >= 2 * sm and batch >= 2 * sm:
            assert s[0] == sm and all(c % sm == 0 for c in s[1:-1]) and s[-1] >= min(sm, batch - sm), (batch, cap, sm, s)


def test_version_and_defaults(lib):
    import ainmf
    assert b"sm_100a" in lib.ainmf_version()
    p = ainmf._capi.default_params(lib)
    assert (p.n_fft, p.hop, p.rank, p.max_iter, p.seed, p.frac_num, p.frac_den) == (1024, 256, 40, 200, 42, 9, 10)
    assert abs(p.tol - 1e-4) < 1e-10 and abs(p.threshold - 1e-4) < 1e-10 and p.col_start == -1 and p.n_outer == 1
    assert [lib.ainmf_padded_rank(k) for k in (1, 32, 33, 40, 64, 65, 128)] == [32, 32, 64, 64, 64, 128, 128]


@pytest.mark.parametrize("N,n_fft,hop", [(441000, 1024, 256), (441000, 2048, 512), (2205, 512, 128),
                                         (158760000, 2048, 512), (1024, 1024, 256), (12345, 256, 64)])
def test_geometry_matches_scipy_restatement(lib, N, n_fft, hop):
    import ainmf
    T, F, ldf = ainmf._capi.stft_geometry(lib, N, n_fft, hop)
    assert (T, F) == (restate.stft_geometry(N, n_fft, hop)[0], n_fft // 2 + 1)
    assert ldf % 4 == 0 and F <= ldf < F + 4


def test_no_cpu_path(lib):
    """Without a CUDA device the handle cannot be created and ops reject CPU tensors: there is no fallback."""
    import torch
    import ainmf
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.ainmf_create(C.byref(h), 0)
    assert rc == ainmf._capi.ERR_NO_DEVICE and b"no CPU path" in lib.ainmf_last_error(None)
    with pytest.raises(NotImplementedError):
        ainmf.ops.stft(torch.zeros(1, 4096), 1024, 256)
    with pytest.raises(RuntimeError):
        ainmf.NMFFairGapInpainter("x.wav", device="cpu")


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "audio-inpainting_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and "emu" not in d:
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|liboracle|oracle[/.]\w", src), os.path.join(d, f)
