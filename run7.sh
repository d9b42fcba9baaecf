python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 > gpurun_out/pytest7.log
python bench.py > gpurun_out/bench7.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench7_ref.log 2>&1
tail -3 gpurun_out/pytest7.log; tail -c 1500 gpurun_out/bench7.log; cat gpurun_out/bench7_ref.log | cut -c1-600
