"""torchrun script: time-frame-sharded restore on N GPUs vs the single-GPU path (rank 0) on the same signal.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_sharded_gpu.py [seconds] [K]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import ainmf
    from ainmf.sharding import TimeShardedInpainter
    import bench
    N = int(seconds * 44100)
    x = bench.c5_signal_device(dev, N)
    tsi = TimeShardedInpainter(dev)
    pl = tsi.plan(N, 2048, 512)
    y, info = tsi.restore(x[pl["x_begin"]:pl["x_end"]].clone(), N, n_fft=2048, hop=512, rank=K, max_iter=40, tol=1e-4, seed=0)
    transport = int(ainmf._lib.lib().ainmf_comm_transport(tsi.h))
    want_transport = 1 if os.environ.get("AINMF_PEER_EXCHANGE") == "0" else 2
    parts = [None] * world
    dist.all_gather_object(parts, (pl["y_begin"], pl["y_end"], y.cpu().numpy(), int(info["n_bad"][0]), int(info["n_iter"][0]),
                                   float(info["err"][0])))
    ok = True
    if rank == 0:
        full = np.zeros(N, np.float32)
        for yb, ye, yy, nb, ni, er in parts:
            full[yb:ye] = yy
        ys, idx, nbs, W, H, err, nit = ainmf.ops.nmf_inpaint(x[None], 2048, 512, K, 40, 1e-4, 0, 1e-4, 9, 10, -1, -1, 1, None, None)
        ys = ys[0].cpu().numpy()
        num = float(np.sum(ys.astype(np.float64) ** 2))
        den = float(np.sum((ys.astype(np.float64) - full) ** 2))
        snr = 10 * np.log10(num / (den + 1e-10))
        rel = abs(parts[0][5] - float(err[0])) / float(err[0])
        print(f"world {world}: N={N} K={K} n_bad {parts[0][3]} vs {int(nbs[0])}, n_iter {parts[0][4]} vs {int(nit[0])}, "
              f"objective rel diff {rel:.2e}, stitched-vs-single SNR {snr:.1f} dB, transport {transport}")
        ok = transport == want_transport and parts[0][3] == int(nbs[0]) and parts[0][4] == int(nit[0]) and rel < 1e-4 and snr > 60
        print("SHARDED_OK" if ok else "SHARDED_FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
