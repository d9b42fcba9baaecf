mkdir -p gpurun_out
python -m pytest tests/test_fd2d_gpu.py tests/test_mc_gpu.py -m gpu -q --timeout 900 2>&1 | tail -5 > gpurun_out/pytest4.log
CMD="python bench.py --nt 200 --steps 1 --warmup 3 --no-cpu-baseline --no-track-a"
$CMD --stream 8,4 > gpurun_out/plain_s.log 2>&1 && ncu --set full --cache-control none --clock-control none --import-source on -k regex:fd2d_st -s 290 -c 12 -o gpurun_out/prof_stream_warm $CMD --stream 8,4 > gpurun_out/ncu_s.log 2>&1
$CMD --tile 16,2 > gpurun_out/plain_t.log 2>&1 && ncu --set full --cache-control none --clock-control none --import-source on -k regex:fd2d_st -s 290 -c 12 -o gpurun_out/prof_tile_warm $CMD --tile 16,2 > gpurun_out/ncu_t.log 2>&1
tail -3 gpurun_out/pytest4.log
