# pileup profile: timing, then ncu source counters of the 3rd pileup_tile launch
python tools/pileup_check.py 12.5 30 5
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section SpeedOfLight --clock-control none --import-source on -k regex:pileup_tile -s 2 -c 1 -f -o gpurun_out/pileup_src python tools/pileup_check.py 12.5 30 3 > gpurun_out/pileup_ncu.log 2>&1
tail -2 gpurun_out/pileup_ncu.log
