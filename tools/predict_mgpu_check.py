#!/usr/bin/env python
"""Multi-GPU drop-in check: `python -m nanosnp_b200.predict` on N GPUs (torchrun) must write the same VCF bytes as on one.

    python tools/predict_mgpu_check.py prepare /tmp/mg        # three synthetic contigs: FASTA + .fai + *.reads.npz
    python -m nanosnp_b200.predict ... -output /tmp/mg/one.vcf
    python -m torch.distributed.run --nproc-per-node 2 ... -m nanosnp_b200.predict ... -output /tmp/mg/two.vcf
    python tools/predict_mgpu_check.py compare /tmp/mg/one.vcf /tmp/mg/two.vcf
(tools/predict_mgpu_check.sh runs the four steps.)"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def prepare(out):
    import numpy as np
    from nanosnp_b200.dataset import save_reads_npz
    from nanosnp_b200.synth import SynthConfig, generate_host
    out = Path(out); (out / "data").mkdir(parents=True, exist_ok=True)
    seqs = {}
    for i, (name, L) in enumerate([("ctgA", 420_000), ("ctgB", 150_000), ("ctgC", 260_000)]):
        cfg = SynthConfig(contig_len=L, coverage=25.0, contig=name, seed_ref=31 + i, seed_var=41 + i, seed_reads=51 + i, len_median=4000, len_min=300)
        ref, reads = generate_host(cfg)
        seqs[name] = ref
        save_reads_npz(str(out / "data" / f"{name}.reads.npz"), reads, name, L)
    with open(out / "ref.fa", "wb") as f, open(str(out / "ref.fa") + ".fai", "w") as fai:      # FASTA + .fai, 60 columns
        off = 0
        for name, seq in seqs.items():
            hdr = f">{name}\n".encode(); f.write(hdr); off += len(hdr)
            L = len(seq); fai.write(f"{name}\t{L}\t{off}\t60\t61\n")
            b = bytes(seq)
            body = b"".join(b[i:i + 60] + b"\n" for i in range(0, L, 60))
            f.write(body); off += len(body)
    print("prepared", out)


def compare(a, b):
    x, y = open(a, "rb").read(), open(b, "rb").read()
    n = sum(1 for l in x.splitlines() if not l.startswith(b"#"))
    print(f"{a}: {len(x)} bytes, {n} records; {b}: {len(y)} bytes -> {'IDENTICAL' if x == y else 'DIFFERENT'}")
    sys.exit(0 if x == y and n > 1000 else 1)


if __name__ == "__main__":
    {"prepare": prepare, "compare": compare}[sys.argv[1]](*sys.argv[2:])
