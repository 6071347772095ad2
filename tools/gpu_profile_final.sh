ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 32 --csv --log-file gpurun_out/launches_final.csv python tools/profile_run.py --chunks 8 > gpurun_out/launches_final.log 2>&1; tail -1 gpurun_out/launches_final.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:k1_mix -s 2 -c 1 -o gpurun_out/k1_final -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_sync -s 2 -c 1 -o gpurun_out/k2_final -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k2a_ -s 2 -c 1 -o gpurun_out/k2a_final -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k3_burst -s 2 -c 1 -o gpurun_out/k3_final -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
ls -la gpurun_out/*final*
