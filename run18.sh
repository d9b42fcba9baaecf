timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 tools/slab_check.py 256 100 > gpurun_out/slab_n4.log 2>&1; echo rc=$?; grep -E "^rank|SLAB" gpurun_out/slab_n4.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n4.log 2>&1; echo rc=$?
python - <<'PY'
import json
for l in open("gpurun_out/bench_n4.log"):
    if l.startswith("{"):
        d=json.loads(l); print("N=4 value", d["value"], "shots/s", d["shots_per_s"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"])
PY
