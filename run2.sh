mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/pytest2.log
python tools/l2_bw.py > gpurun_out/l2_bw.log 2>&1
CMD="python bench.py --nt 300 --steps 1 --warmup 3 --no-cpu-baseline --no-track-a"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fd2d_step -s 290 -c 20 -o gpurun_out/prof_fd2d_r1 $CMD > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/pytest2.log; cat gpurun_out/l2_bw.log
