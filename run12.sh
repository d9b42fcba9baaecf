python -m pytest tests/test_fd2d_gpu.py -m gpu -q --timeout 900 -k "variants or gradient" 2>&1 | tail -3
for pers in 1 0; do for v in "--tile 32,4" "--tile 16,2" "--tb2 32" "--tb2 24"; do echo "persist=$pers $v"; FWI_L2_PERSIST=$pers python bench.py --nt 3000 --steps 4 --warmup 3 $v --no-cpu-baseline --no-track-a 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   value %.1f  e2e %.1f  avg_launch_us %.2f launches %d'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['gpu_launches']))
    elif 'rror' in l: print(l.strip()[:300])
"; done; done
