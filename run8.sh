python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py > gpurun_out/dist_check.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench8_n2.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1 --warmup 0 --impl reference > gpurun_out/bench8_n2_ref.log 2>&1
tail -3 gpurun_out/dist_check.log; tail -c 900 gpurun_out/bench8_n2.log; tail -c 300 gpurun_out/bench8_n2_ref.log
