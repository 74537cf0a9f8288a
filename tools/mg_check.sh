TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29541 tests/run_sharded_gpu.py 60 128 > gpurun_out/mg2_sharded_peer.log 2>&1; echo "sharded peer rc=$?"; grep -h "world\|SHARDED" gpurun_out/mg2_sharded_peer.log
AINMF_PEER_EXCHANGE=0 $TR --master-port 29542 tests/run_sharded_gpu.py 60 128 > gpurun_out/mg2_sharded_nccl.log 2>&1; echo "sharded nccl rc=$?"; grep -h "world\|SHARDED" gpurun_out/mg2_sharded_nccl.log
$TR --master-port 29543 bench.py --gpus 2 --workload c5 --no-cpu-baseline > gpurun_out/mg2_c5_peer.json 2> gpurun_out/mg2_c5_peer.err; echo "c5 peer rc=$?"
AINMF_PEER_EXCHANGE=0 $TR --master-port 29544 bench.py --gpus 2 --workload c5 --no-cpu-baseline > gpurun_out/mg2_c5_nccl.json 2> gpurun_out/mg2_c5_nccl.err; echo "c5 nccl rc=$?"
tail -c 400 gpurun_out/mg2_c5_peer.err
