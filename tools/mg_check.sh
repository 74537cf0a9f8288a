#!/bin/bash
# Multi-GPU check on one box: time-sharded parity (both transports) and the c5 bench line.  Usage: bash tools/mg_check.sh N [quick]
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$2" != "quick" ]; then
$TR --master-port 29541 tests/run_sharded_gpu.py 60 128 > gpurun_out/mg${N}_sharded_peer.log 2>&1; echo "sharded peer rc=$?"; grep -h "world\|SHARDED" gpurun_out/mg${N}_sharded_peer.log
AINMF_PEER_EXCHANGE=0 $TR --master-port 29542 tests/run_sharded_gpu.py 60 128 > gpurun_out/mg${N}_sharded_nccl.log 2>&1; echo "sharded nccl rc=$?"; grep -h "world\|SHARDED" gpurun_out/mg${N}_sharded_nccl.log
fi
$TR --master-port 29543 bench.py --gpus $N --workload c5 --no-cpu-baseline > gpurun_out/mg${N}_c5_peer.json 2> gpurun_out/mg${N}_c5_peer.err; echo "c5 peer rc=$?"
AINMF_PEER_EXCHANGE=0 $TR --master-port 29544 bench.py --gpus $N --workload c5 --no-cpu-baseline > gpurun_out/mg${N}_c5_nccl.json 2> gpurun_out/mg${N}_c5_nccl.err; echo "c5 nccl rc=$?"
tail -c 300 gpurun_out/mg${N}_c5_peer.err
