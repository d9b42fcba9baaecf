set +e
O=gpurun_out
CMD="python bench.py --nt 300 --steps 1 --warmup 3 --no-cpu-baseline --no-track-a"
$CMD > $O/plain20.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_r1c.csv $CMD > $O/ncu20a.log 2>&1
python tools/summarize_ncu.py launches $O/launches_r1c.csv > $O/r1_launches.txt; rm -f $O/launches_r1c.csv
# the bench's warm-up forward (300 plain launches) comes first, then 3 warm-up gradients (300 fwd-save + 300 adj each)
$CMD > $O/plain20b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fd2d_step -s 896 -c 8 -o $O/prof_fd2d_cold $CMD > $O/ncu20b.log 2>&1
ncu -i $O/prof_fd2d_cold.ncu-rep --page raw --csv > /tmp/a.csv 2>/dev/null; python tools/summarize_ncu.py raw /tmp/a.csv > $O/r1_fd2d_cold.txt
$CMD > $O/plain20c.log 2>&1 && ncu --set full --cache-control none --clock-control none -k regex:fd2d_step -s 896 -c 8 -o $O/prof_fd2d_warm $CMD > $O/ncu20c.log 2>&1
ncu -i $O/prof_fd2d_warm.ncu-rep --page raw --csv > /tmp/b.csv 2>/dev/null; python tools/summarize_ncu.py raw /tmp/b.csv > $O/r1_fd2d_warm.txt; rm -f $O/prof_fd2d_warm.ncu-rep
python tools/step_bench3d.py 384 > $O/plain20d.log 2>&1 && ncu --set full --clock-control none -k regex:fd3d -s 20 -c 2 -o $O/prof_fd3d python tools/step_bench3d.py 384 > $O/ncu20d.log 2>&1
ncu -i $O/prof_fd3d.ncu-rep --page raw --csv > /tmp/c.csv 2>/dev/null; python tools/summarize_ncu.py raw /tmp/c.csv > $O/r1_fd3d.txt; rm -f $O/prof_fd3d.ncu-rep
python tools/mc_bench.py 2000000 > $O/r1_mc_bench.txt 2>&1 && ncu --set full --clock-control none -k regex:mc_eval -s 2 -c 1 -o $O/prof_mc python tools/mc_bench.py 2000000 > $O/ncu20e.log 2>&1
ncu -i $O/prof_mc.ncu-rep --page raw --csv > /tmp/d.csv 2>/dev/null; python tools/summarize_ncu.py raw /tmp/d.csv > $O/r1_mc.txt; rm -f $O/prof_mc.ncu-rep
ls -la $O | head -30; du -sh $O
