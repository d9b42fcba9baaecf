for i in 1 2 3; do for v in "--tile 32,4" "--tb2 24"; do echo "run $i $v"; python bench.py $v --no-cpu-baseline --no-track-a 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   value %.1f  e2e %.1f  avg_launch_us %.2f  clocks %s'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['clocks']))
    elif 'rror' in l: print(l.strip()[:300])
"; done; done
