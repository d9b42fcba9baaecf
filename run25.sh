python tools/step_bench.py tile:24,4 tile:24,3 tile:24,2 tile:48,4 tile:32,4 2>&1 | grep -E "1000 x|^tile|graphs" | grep -B1 "1000 x" | grep -v "^--"
for t in 24,4 24,3 32,4; do echo "bench tile $t"; python bench.py --tile $t --no-cpu-baseline --no-track-a 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   value %.1f  e2e %.1f  avg_launch_us %.2f'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_us']))
    elif 'rror' in l: print(l.strip()[:300])
"; done
