python -m pytest tests -x -q -m gpu --timeout 900 2>&1 | tail -4 > gpurun_out/pytest19.log; cat gpurun_out/pytest19.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CMD="python bench.py --nt 300 --steps 1 --warmup 3 --no-cpu-baseline --no-track-a"
$CMD > gpurun_out/plain19.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu19a.log 2>&1
$CMD > gpurun_out/plain19b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fd2d_step -s 890 -c 20 -o gpurun_out/prof_fd2d_r1c $CMD > gpurun_out/ncu19b.log 2>&1
$CMD > gpurun_out/plain19c.log 2>&1 && ncu --set full --cache-control none --clock-control none -k regex:fd2d_step -s 890 -c 20 -o gpurun_out/prof_fd2d_r1c_warm $CMD > gpurun_out/ncu19c.log 2>&1
python tools/step_bench3d.py 384 > gpurun_out/plain19d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fd3d -s 20 -c 2 -o gpurun_out/prof_fd3d_r1c python tools/step_bench3d.py 384 > gpurun_out/ncu19d.log 2>&1
python tools/mc_bench.py 2000000 > gpurun_out/mc_bench_r1c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mc_eval -s 2 -c 1 -o gpurun_out/prof_mc_r1c python tools/mc_bench.py 2000000 > gpurun_out/ncu19e.log 2>&1
tail -1 gpurun_out/ncu19a.log gpurun_out/ncu19b.log gpurun_out/ncu19c.log gpurun_out/ncu19d.log gpurun_out/ncu19e.log
