#!/usr/bin/env python
"""Stand-alone pileup timing on one synthetic region (for ncu).  usage: tools/pileup_check.py [mb] [coverage] [reps] [ins,del,sub rates]
(default rates: the bench workload's 0.03,0.04,0.03 = one CIGAR op per 7.5 aligned bases; e.g. 0.004,0.006,0.008 for a Q20-like profile)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nanosnp_b200.pipeline import PileupEngine
from nanosnp_b200.synth import SynthConfig, generate_device

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 12.5
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
rates = [float(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else None
cfg = SynthConfig(contig_len=int(mb * 1e6), coverage=cov, seed_ref=1000, seed_var=1001, seed_reads=1002)
if rates:
    cfg.ins_rate, cfg.del_rate, cfg.sub_rate = rates
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
alg = 0.25 * reads.n_bases + 2 * reads.n_cigar + 23 * reads.n_reads + 74 * cfg.contig_len      # DESIGN 4.1 algorithmic bytes (uint16 CIGAR words as shipped: 4 -> 2)
print(f"ops/base={reads.n_cigar / max(1, reads.n_bases):.4f} algorithmic GB/s at the best run: {alg / (min(ts) * 1e-3) / 1e9:.0f}")
print(f"pileup {mb} Mb {cov}x: reads={reads.n_reads} cigar={reads.n_cigar} ms={['%.3f' % t for t in ts]} checksum={int(counts.sum().item())} gate={int((flags & 2).ne(0).sum().item())}")
