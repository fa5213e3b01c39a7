# multi-GPU checks of round 2 (gpurun --gpus N): the bench under torchrun (pv_allgather_photons = NCCL inside libpv.so) and the
# drop-in binary with PV_DEVICES (pv_comm_init_all + pv_broadcast_photons)
cd /root/repo
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_mg_n$N.log 2>&1
tail -1 gpurun_out/r02_mg_n$N.log | cut -c1-400
grep -o '"allgather": {[^}]*}' gpurun_out/r02_mg_n$N.log; grep -o '"frame": {[^}]*}' gpurun_out/r02_mg_n$N.log
T=$(mktemp -d); cd $T
DEVS=$(python -c "print(','.join(str(i) for i in range($N)))")
/root/repo/baseline/_ref/pbrt_b200 --quiet /root/repo/tests/scenes/cornell_e2e.pbrt 2> one.err; mv cornell_e2e.pfm one.pfm
PV_DEVICES=$DEVS /root/repo/baseline/_ref/pbrt_b200 --quiet /root/repo/tests/scenes/cornell_e2e.pbrt 2> many.err; mv cornell_e2e.pfm many.pfm
cmp one.pfm many.pfm && echo "PV_DEVICES=$DEVS: image identical to one device"
grep "\[pv\]" one.err many.err
( echo "== one device"; cat one.err; echo "== PV_DEVICES=$DEVS"; cat many.err ) > /root/repo/gpurun_out/r02_pv_devices_n$N.log
