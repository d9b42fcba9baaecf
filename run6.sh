python -m pytest tests/test_fd2d_gpu.py -m gpu -q --timeout 900 2>&1 | tail -8 > gpurun_out/pytest6.log
python tools/step_bench.py > gpurun_out/step_bench2.log 2>&1
python bench.py --no-cpu-baseline --no-track-a > gpurun_out/bench6.log 2>&1
tail -4 gpurun_out/pytest6.log; cat gpurun_out/step_bench2.log; python -c "
import json
for l in open('gpurun_out/bench6.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['gpu_launches'])
    else: print(l.strip()[:200])
"
