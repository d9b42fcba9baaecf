python -m pytest tests/test_fd2d_gpu.py -m gpu -q --timeout 900 2>&1 | tail -4
for i in 1 2; do python bench.py --no-cpu-baseline --no-track-a 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   value %.1f  e2e %.1f  avg_launch_us %.2f frac %.3f'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac']))
    elif 'rror' in l: print(l.strip()[:300])
"; done
