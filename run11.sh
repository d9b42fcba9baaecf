python -m pytest tests/test_fd2d_gpu.py tests/test_fd3d_gpu.py -m gpu -q --timeout 900 2>&1 | tail -12
python tools/step_bench.py tb2:32 tb2:16 tb2:24 tb2:56 2>&1 | grep -v "graphs = False" 
for c in 32 24 56; do echo "tb2 $c"; python bench.py --nt 3000 --steps 4 --warmup 3 --tb2 $c --no-cpu-baseline --no-track-a 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   value %.1f  e2e %.1f  avg_launch_us %.2f launches %d'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['gpu_launches']))
    elif 'rror' in l: print(l.strip()[:300])
"; done
