cd /root/repo
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload config5 --steps 3 --warmup 3 --shoot-photons 0 > gpurun_out/cfg5_n$N.log 2>&1
tail -2 gpurun_out/cfg5_n$N.log | cut -c1-1500
