# round-2 profiles: launch list of the bench command (25 Mb, shares), full sets of the s1 / record / text kernels
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 260 --csv --log-file gpurun_out/r02_launches_25mb.csv \
    python bench.py --contig-mb 25 --steps 2 --warmup 3 --no-cpu-baseline --no-selfcheck > gpurun_out/r02_launch_bench.json 2> gpurun_out/r02_launch_bench.err
ncu --set full --clock-control none --import-source on -k regex:"pileup_tile|read_scan|cap_kernel" -s 6 -c 3 -f -o gpurun_out/r02_pileup python tools/pileup_check.py 12.5 30 4 > gpurun_out/r02_pileup_ncu.log 2>&1
ncu --set full --clock-control none -k regex:"vcf_text|block_scan|site_record|select_|lstm0_pair2" -s 0 -c 12 -f -o gpurun_out/r02_text python bench.py --contig-mb 12.5 --steps 1 --warmup 3 --no-cpu-baseline --no-selfcheck > gpurun_out/r02_text_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
