"""Multi-GPU use of the path, one process per GPU (torch.distributed for the plumbing).

Two partitions (SURVEY 8e / BASELINE configs[3], configs[4]):
  * clip batches: clips are independent -- `shard_clips` gives each rank its slice, no collective on the data path;
  * one long signal split by time frames: each rank owns a slice of H and V and a replicated W; per iteration the
    F*K + K*K partial sums of the W half-step are all-reduced over NVLink by NCCL inside libainmf.so
    (`TimeShardedInpainter`).  torch.distributed only carries the 128-byte NCCL id and the final gather.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _capi, _lib


def shard_clips(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of the clips rank `rank` of `world` processes (contiguous, sizes differ by at most one)."""
    base, rem = divmod(n_clips, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_plan(n_samples: int, n_fft: int, hop: int, rank: int, world: int) -> dict:
    """Frames / samples of rank `rank` in the time-frame split (ainmf_shard_plan)."""
    L = _lib.lib()
    tb, te = C.c_int32(), C.c_int32()
    xb, xe, yb, ye = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    rc = L.ainmf_shard_plan(n_samples, n_fft, hop, rank, world, C.byref(tb), C.byref(te), C.byref(xb), C.byref(xe),
                            C.byref(yb), C.byref(ye))
    if rc:
        raise _capi.AinmfError(rc, f"cannot split N={n_samples} n_fft={n_fft} hop={hop} over {world} ranks")
    return dict(t_begin=tb.value, t_end=te.value, x_begin=xb.value, x_end=xe.value, y_begin=yb.value, y_end=ye.value)


class TimeShardedInpainter:
    """One long signal restored by all ranks of the default process group (backend nccl, one GPU per rank)."""

    def __init__(self, device: torch.device | None = None, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("ainmf runs on CUDA devices only (no CPU fallback)")
        self.dev = self.device.index
        L = _lib.lib()
        self.h = _lib.handle(self.dev)
        # NCCL communicator of our own for the per-iteration all-reduce: rank 0 makes the id, torch broadcasts it
        idt = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            _lib.check(L.ainmf_comm_unique_id(buf), self.dev)
            idt = torch.tensor(list(buf), dtype=torch.uint8)
        idt = idt.to(self.device)
        dist.broadcast(idt, src=0, group=group)
        raw = (C.c_uint8 * 128)(*idt.cpu().tolist())
        with torch.cuda.device(self.dev):
            _lib.check(L.ainmf_comm_init(self.h, raw, self.rank, self.world), self.dev)
        self._ws = None

    def plan(self, n_samples, n_fft, hop):
        return shard_plan(n_samples, n_fft, hop, self.rank, self.world)

    def restore(self, x_local: torch.Tensor, n_samples: int, *, n_fft=2048, hop=512, rank=128, max_iter=200, tol=1e-4,
                seed=0, threshold=1e-4, frac=(9, 10)):
        """x_local: this rank's slice x[x_begin:x_end] (float32, CUDA).  Returns (y_local, info): y_local is
        y[y_begin:y_end]; info holds n_bad (global), n_iter, err (global objective), W (replicated), H_local."""
        L = _lib.lib()
        pl = self.plan(n_samples, n_fft, hop)
        if x_local.dtype != torch.float32 or not x_local.is_cuda or x_local.numel() != pl["x_end"] - pl["x_begin"]:
            raise RuntimeError("x_local must be the float32 CUDA slice x[x_begin:x_end] of ainmf.sharding.shard_plan")
        x_local = x_local.contiguous()
        p = _capi.default_params(L, batch=1, n_samples=n_samples, n_fft=n_fft, hop=hop, rank=rank, max_iter=max_iter,
                                 tol=tol, seed=seed & 0xFFFFFFFF, threshold=threshold, frac_num=frac[0], frac_den=frac[1])
        with torch.cuda.device(self.dev):
            nbytes = L.ainmf_sharded_workspace_bytes(self.h, C.byref(p))
            if nbytes == 0:
                _lib.check(_capi.ERR_INVALID, self.dev)
            if self._ws is None or self._ws.numel() < nbytes:
                self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            F = n_fft // 2 + 1
            Tl = pl["t_end"] - pl["t_begin"]
            y = torch.empty(pl["y_end"] - pl["y_begin"], dtype=torch.float32, device=self.device)
            nb = torch.zeros(1, dtype=torch.int32, device=self.device)
            nit = torch.zeros(1, dtype=torch.int32, device=self.device)
            err = torch.zeros(1, dtype=torch.float32, device=self.device)
            W = torch.zeros((F, rank), dtype=torch.float32, device=self.device)
            Hl = torch.zeros((rank, Tl), dtype=torch.float32, device=self.device)
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            vp = lambda t: C.c_void_p(t.data_ptr())
            _lib.check(L.ainmf_inpaint_sharded(self.h, C.byref(p), vp(x_local), vp(y), vp(nb), vp(W), vp(Hl), vp(err),
                                               vp(nit), vp(self._ws), self._ws.numel(), stream), self.dev)
        return y, dict(n_bad=nb, n_iter=nit, err=err, W=W, H_local=Hl, plan=pl)
