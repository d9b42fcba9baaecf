mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -40 > gpurun_out/pytest1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke1.log 2>&1
python bench.py --nt 1000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1
for t in 32,4 32,8 16,4 64,8 16,2 64,4; do python bench.py --nt 1000 --steps 3 --warmup 3 --tile $t --no-cpu-baseline --no-track-a > gpurun_out/bench_tile_$t.log 2>&1; done
python bench.py > gpurun_out/bench_full.log 2>&1
tail -3 gpurun_out/pytest1.log; cat gpurun_out/smoke1.log | tail -3; cat gpurun_out/bench_small.log | tail -2
