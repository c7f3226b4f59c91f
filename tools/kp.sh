# pileup: parity + timing of tile sizes
NSNP_PILEUP_TILE=2048 python -m pytest tests/test_gpu_s1.py tests/test_gpu_scale.py -x -q -m gpu 2>&1 | tail -3
python tools/pileup_check.py 12.5 30 5
NSNP_PILEUP_TILE=2048 python tools/pileup_check.py 12.5 30 5
NSNP_PILEUP_TILE=2048 NSNP_PILEUP_VARIANT=1 python tools/pileup_check.py 12.5 30 5
