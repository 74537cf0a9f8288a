"""Time-frame sharding on CPU: world_size-2 (and 3) gloo process groups drive the REAL sharded algorithm of
csrc/sharded.inc -- compiled for the host by the emulator harness -- with the two collectives supplied as
callbacks over torch.distributed.  The stitched result must equal the single-rank result and the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _signal(N, sr=8000):
    rng = np.random.default_rng(5)
    t = np.arange(N) / sr
    x = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.3 * np.sin(2 * np.pi * 1230 * t) + 0.05 * rng.standard_normal(N)).astype(np.float32)
    x[2500:3300] = 0
    x[7000:7700] = 0
    return x / np.abs(x).max()


def _worker(rank, world, port, N, kw, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import emu_harness as E
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib, h, capi = E.lib(), E.handle(), E.capi
    np_dt = {0: np.float32, 1: np.float64, 2: np.int32}

    def allreduce(user, buf, count, dtype, op, stream):
        a = np.ctypeslib.as_array(C.cast(buf, C.POINTER({0: C.c_float, 1: C.c_double, 2: C.c_int32}[dtype])), shape=(count,))
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX)
        return 0

    def sendrecv(user, sbuf, speer, rbuf, rpeer, n, stream):
        reqs = []
        if speer >= 0:
            s = torch.from_numpy(np.ctypeslib.as_array(C.cast(sbuf, C.POINTER(C.c_float)), shape=(n,)).copy())
            reqs.append(dist.isend(s, dst=speer))
        if rpeer >= 0:
            r = torch.from_numpy(np.ctypeslib.as_array(C.cast(rbuf, C.POINTER(C.c_float)), shape=(n,)))
            reqs.append(dist.irecv(r, src=rpeer))
        for q in reqs:
            q.wait()
        return 0

    ar, sr = capi.ALLREDUCE_FN(allreduce), capi.SENDRECV_FN(sendrecv)
    E.check(lib.ainmf_comm_set_callbacks(h, rank, world, ar, sr, None))
    x = _signal(N)
    p = capi.default_params(lib, batch=1, n_samples=N, **kw)
    tb, te = C.c_int32(), C.c_int32()
    xb, xe, yb, ye = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    E.check(lib.ainmf_shard_plan(N, p.n_fft, p.hop, rank, world, C.byref(tb), C.byref(te), C.byref(xb), C.byref(xe),
                                 C.byref(yb), C.byref(ye)))
    xl = np.ascontiguousarray(x[xb.value:xe.value])
    nbytes = lib.ainmf_sharded_workspace_bytes(h, C.byref(p))
    assert nbytes > 0
    ws = np.zeros(nbytes + 256, np.uint8)
    off = (-ws.ctypes.data) % 256
    F, K, Tl = p.n_fft // 2 + 1, p.rank, te.value - tb.value
    y = np.zeros(ye.value - yb.value, np.float32)
    nb, nit, err = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.float32)
    W, Hl = np.zeros((F, K), np.float32), np.zeros((K, Tl), np.float32)
    E.check(lib.ainmf_inpaint_sharded(h, C.byref(p), E.ptr(xl), E.ptr(y), E.ptr(nb), E.ptr(W), E.ptr(Hl), E.ptr(err),
                                      E.ptr(nit), C.c_void_p(ws.ctypes.data + off), nbytes, None))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), y=y, yb=yb.value, ye=ye.value, tb=tb.value, te=te.value, nb=nb,
             nit=nit, err=err, W=W, Hl=Hl)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,max_iter,tol", [(2, 6, 1e-4), (3, 6, 1e-4), (2, 40, 5e-2)])
def test_time_sharded_equals_single_rank(world, max_iter, tol, tmp_path):
    """The third case stops early (tol = 5e-2): the H-side violation of an iteration travels with the next iteration's
    exchange and the rule is evaluated one exchange late, before anything is updated -- n_iter and the factors must
    still be those of the iteration at which sklearn stops."""
    import torch.multiprocessing as mp
    import emu_harness as E
    from oracle import libcalls
    N = 10000
    kw = dict(n_fft=128, hop=32, rank=8, max_iter=max_iter, tol=tol, seed=42)
    port = 29500 + (os.getpid() % 1000) + world
    mp.spawn(_worker, args=(world, port, N, kw, str(tmp_path)), nprocs=world, join=True)
    x = _signal(N)
    single = E.inpaint(x, **kw)
    yo, st = libcalls.restore_columns(x, 8000, n_fft=128, hop=32, K=8, seed=42, max_iter=max_iter, tol=tol, return_all=True)
    if max_iter > 6:
        assert 1 < st["n_iter"] < max_iter                               # the case is an early stop
    y = np.zeros(N, np.float32)
    cover = np.zeros(N, np.int32)
    H = np.zeros((8, single["H"].shape[2]), np.float32)
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        y[d["yb"]:d["ye"]] = d["y"]
        cover[d["yb"]:d["ye"]] += 1
        H[:, d["tb"]:d["te"]] = d["Hl"]
        assert d["nb"][0] == single["n_bad"][0] == len(st["bad"])          # global count on every rank
        assert d["nit"][0] == single["n_iter"][0] == st["n_iter"]
        assert abs(d["err"][0] - st["err"]) < 1e-5 * st["err"]            # global objective on every rank
        assert np.abs(d["W"] - single["W"][0]).max() < 1e-4 * np.abs(single["W"]).max()   # W replicated
    assert np.all(cover == 1)                                             # the y slices tile [0, N)
    assert np.abs(H - single["H"][0]).max() < 1e-4 * np.abs(single["H"]).max()
    assert libcalls.snr_db(single["y"][0], y) > 100                        # stitched == single rank
    assert libcalls.snr_db(yo, y) > 90                                     # == oracle


def test_shard_plan_tiles_frames_and_samples():
    import emu_harness as E
    lib = E.lib()
    for N, n_fft, hop, world in [(158760000, 2048, 512, 8), (441000, 1024, 256, 4), (10000, 128, 32, 3), (441000, 2048, 512, 2)]:
        T, _, _ = E.capi.stft_geometry(lib, N, n_fft, hop)
        prev_t, prev_y = 0, 0
        for r in range(world):
            tb, te = C.c_int32(), C.c_int32()
            xb, xe, yb, ye = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
            assert lib.ainmf_shard_plan(N, n_fft, hop, r, world, C.byref(tb), C.byref(te), C.byref(xb), C.byref(xe),
                                        C.byref(yb), C.byref(ye)) == 0
            assert tb.value == prev_t and yb.value == prev_y and te.value > tb.value
            assert 0 <= xb.value <= yb.value and ye.value <= xe.value <= N
            # halo: the mask window and frame support of every owned frame lie inside [x_begin, x_end) or outside [0, N)
            assert xb.value <= max(0, tb.value * hop - n_fft // 2) and xe.value >= min(N, (te.value - 1) * hop + n_fft // 2)
            prev_t, prev_y = te.value, ye.value
        assert prev_t == T and prev_y == N
    assert lib.ainmf_shard_plan(2000, 1024, 256, 0, 8, None, None, None, None, None, None) != 0   # too few frames
