#!/usr/bin/env python
"""Inflate + decode throughput of the streaming BAM reader (csrc/bam_stream.cu) on this host.
usage: tools/bam_bench.py [contig_mb=12.5] [coverage=30] [threads=0 (all cores)]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanosnp_b200.bam import BamReader, write_bam
from nanosnp_b200.synth import SynthConfig, generate_host

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 12.5
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = SynthConfig(contig_len=int(mb * 1e6), coverage=cov, contig="ctg1", seed_ref=1000, seed_var=2000, seed_reads=3000)
t = time.time(); ref, reads = generate_host(cfg); t_gen = time.time() - t
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    path = os.path.join(d, "b.bam")
    t = time.time(); write_bam(path, [("ctg1", cfg.contig_len)], {"ctg1": reads}, index=True); t_w = time.time() - t
    size = os.path.getsize(path)
    for rep in range(3):
        t = time.time()
        with BamReader(path, threads) as r:
            got = [(name, rd.n_reads, rd.n_cigar) for _, name, rd in r.contigs()]
            raw = r.inflated_bytes
        dt = time.time() - t
        print(f"decode pass {rep}: {dt:.3f} s  compressed {size / 1e6:.0f} MB ({size / dt / 1e9:.2f} GB/s)  inflated {raw / 1e6:.0f} MB ({raw / dt / 1e9:.2f} GB/s)  "
              f"reads {got[0][1]} ops {got[0][2]}  threads {threads or os.cpu_count()}")
    t = time.time()
    with BamReader(path, threads) as r:
        part = r.fetch(0, cfg.contig_len // 2, cfg.contig_len // 2 + 1_000_000)
        raw = r.inflated_bytes
    print(f"fetch 1 Mb window through the .bai: {time.time() - t:.3f} s, {part.n_reads} reads, {raw / 1e6:.0f} MB inflated")
print(f"(generate {t_gen:.1f} s, write {t_w:.1f} s)")
