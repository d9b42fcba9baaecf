for t in 16,2 32,4 32,8 64,8 64,4 16,4; do echo "tile $t"; python bench.py --nt 3000 --steps 4 --warmup 3 --tile $t --no-cpu-baseline --no-track-a 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   value %.1f  e2e %.1f  avg_launch_us %.2f'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_us']))
    elif 'rror' in l: print(l.strip()[:200])
"; done
