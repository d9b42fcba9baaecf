mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/pytest3.log
for s in 8,4 4,8 6,5 8,3 12,3; do python bench.py --nt 2000 --steps 3 --warmup 3 --stream $s --no-cpu-baseline --no-track-a > gpurun_out/bench_stream_$s.log 2>&1; done
python bench.py --nt 2000 --steps 3 --warmup 3 --tile 16,2 --no-cpu-baseline --no-track-a > gpurun_out/bench_tile2_16,2.log 2>&1
tail -5 gpurun_out/pytest3.log; grep -h -o '"value": [0-9.]*' gpurun_out/bench_stream_*.log gpurun_out/bench_tile2*.log
