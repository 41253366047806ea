ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_608.csv python scripts/step_for_ncu.py 32 608 0.01 > gpurun_out/ncu608.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_608.csv | head -24
