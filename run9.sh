CMD="python bench.py --nt 300 --steps 1 --warmup 3 --no-cpu-baseline --no-track-a"
$CMD > gpurun_out/plain9.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu9a.log 2>&1
$CMD > gpurun_out/plain9b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fd2d_step -s 290 -c 20 -o gpurun_out/prof_fd2d_r1b $CMD > gpurun_out/ncu9b.log 2>&1
python tools/mc_bench.py > gpurun_out/mc_bench.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mc_eval -s 2 -c 2 -o gpurun_out/prof_mc_r1 python tools/mc_bench.py 2000000 > gpurun_out/ncu9c.log 2>&1
cat gpurun_out/mc_bench.log; tail -2 gpurun_out/ncu9a.log gpurun_out/ncu9b.log gpurun_out/ncu9c.log
