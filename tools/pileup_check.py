#!/usr/bin/env python
"""Stand-alone pileup timing on one synthetic region (for ncu).  usage: tools/pileup_check.py [mb] [coverage] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nanosnp_b200.pipeline import PileupEngine
from nanosnp_b200.synth import SynthConfig, generate_device

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 12.5
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
cfg = SynthConfig(contig_len=int(mb * 1e6), coverage=cov, seed_ref=1000, seed_var=1001, seed_reads=1002)
ref, reads = generate_device(cfg, dev)
eng = PileupEngine(dev)
counts = torch.empty((cfg.contig_len, 18), dtype=torch.int32, device=dev)
flags = torch.empty(cfg.contig_len, dtype=torch.uint8, device=dev)
ts = []
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.pileup_counts(reads, ref, 0, cfg.contig_len, counts, flags); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
eng.check_status()
print(f"pileup {mb} Mb {cov}x: reads={reads.n_reads} cigar={reads.n_cigar} ms={['%.3f' % t for t in ts]} checksum={int(counts.sum().item())} gate={int((flags & 2).ne(0).sum().item())}")
