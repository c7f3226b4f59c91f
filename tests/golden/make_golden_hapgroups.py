#!/usr/bin/env python
"""Generates tests/golden/hapgroups_small.npz by running the REFERENCE's own HaplotypeModel s4 code
(select_hetesnp_homosnp.py, create_pileup_haplotype.py, make_predict_bins.py, write_to_bins.py from /root/reference, imported
unmodified) on a small synthetic HP-tagged read set.  Run in the build container only:

    python tests/golden/make_golden_hapgroups.py

pysam and PyTables are not installed, so two stubs are injected before the import: `pysam` = oracle/pysam_emul.py (a restatement
of htslib's pileup engine: that step is therefore UNPINNED; everything the reference does with the columns is the reference's own
code) and `tables` = a recorder that keeps the arrays write_to_bins.py appends.
"""
import os
import random
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pysam_emul  # noqa: E402

REF = "/root/reference/HaplotypeModel"


# ------------------------------------------------------------------ synthetic input
def make_input(seed=11):
    rng = random.Random(seed)
    contigs = [("chrA", 24000), ("chrB", 14000), ("chrC", 12000)]
    refs, reads, vcf = {}, {}, []
    for ctg, L in contigs:
        ref = "".join(rng.choice("ACGT") for _ in range(L))
        refs[ctg] = ref
        # heterozygous sites every ~250 bp: hap[0] carries ref, hap[1] the alt (or the other way round)
        sites = []
        p = rng.randrange(100, 300)
        while p < L - 100:
            alt = rng.choice([b for b in "ACGT" if b != ref[p]])
            sites.append((p, alt, rng.randrange(2)))
            p += rng.randrange(120, 380)
        haps = []
        for h in range(2):
            s = list(ref)
            for (q, alt, which) in sites:
                if which == h:
                    s[q] = alt
            haps.append("".join(s))
        # pileup VCF rows as PileupModel/predict.py writes them (columns 0,1,3,4,5,9 are what s4 parses)
        for (q, alt, which) in sites:
            r = rng.random()
            gt = "0/1" if r < 0.8 else "1/1" if r < 0.88 else "0/0" if r < 0.94 else "1/2"
            qual = round(rng.uniform(2, 60), 2)
            vcf.append((ctg, q + 1, ref[q], alt, qual, gt))
        for _ in range(len(sites) // 5):                                  # extra homozygous / low-quality rows between the sites
            q = rng.randrange(50, L - 50)
            if any(q == s[0] for s in sites):
                continue
            gt = rng.choice(["1/1", "0/0", "0/1"])
            vcf.append((ctg, q + 1, ref[q], rng.choice("ACGT"), round(rng.uniform(2, 40), 2), gt))
        recs = []
        depth = 22
        n_reads = depth * L // 4000
        starts = sorted(rng.randrange(0, L - 600) for _ in range(n_reads))
        if ctg == "chrA":                                                 # a deep spot for the max_coverage rule
            starts = sorted(starts + [rng.randrange(9000, 9400) for _ in range(30)])
        for i, st in enumerate(starts):
            h = rng.randrange(2)
            span = min(rng.randrange(1500, 6500), L - st)
            cig, seq = [], []
            x = st
            if rng.random() < 0.2:
                k = rng.randrange(5, 40); cig.append((k, "S")); seq.append("".join(rng.choice("ACGT") for _ in range(k)))
            if rng.random() < 0.03:
                cig.append((rng.randrange(1, 4), "I")); seq.append("".join(rng.choice("ACGT") for _ in range(cig[-1][0])))
            end = st + span
            skip_at = rng.randrange(st + 200, end - 200) if (rng.random() < 0.04 and span > 800) else -1
            while x < end:
                run = min(rng.randrange(1, 30), end - x)
                s = list(haps[h][x:x + run])
                for j in range(run):
                    if rng.random() < 0.03:
                        s[j] = rng.choice("ACGT")
                    if rng.random() < 0.00002:
                        s[j] = "N"
                if cig and cig[-1][1] == "M":
                    cig[-1] = (cig[-1][0] + run, "M")
                else:
                    cig.append((run, "M"))
                seq.append("".join(s)); x += run
                if x >= end:
                    break
                r = rng.random()
                if skip_at >= 0 and x >= skip_at:
                    k = min(rng.randrange(20, 80), end - x - 1)
                    if k > 0:
                        cig.append((k, "N")); x += k
                    skip_at = -1
                elif r < 0.25:
                    k = min(rng.randrange(1, 5), end - x - 1)
                    if k > 0:
                        cig.append((k, "D")); x += k
                elif r < 0.5:
                    k = rng.randrange(1, 5)
                    cig.append((k, "I")); seq.append("".join(rng.choice("ACGT") for _ in range(k)))
            if cig[-1][1] != "M":                                         # keep the last reference-consuming op a match
                cig.append((1, "M")); seq.append(haps[h][min(x, L - 1)]); x += 1
            if rng.random() < 0.2:
                k = rng.randrange(5, 40); cig.append((k, "S")); seq.append("".join(rng.choice("ACGT") for _ in range(k)))
            seq = "".join(seq)
            r = rng.random()
            flag = 16 if rng.random() < 0.5 else 0
            if r < 0.03: flag |= 256
            elif r < 0.05: flag |= 1024
            elif r < 0.06: flag |= 4
            elif r < 0.09: flag |= 2048
            elif r < 0.10: flag |= 1                                      # paired, not a proper pair: dropped by ignore_orphans
            elif r < 0.11: flag |= 3                                      # proper pair: kept
            elif r < 0.115: flag |= 512
            hp = (h + 1) if rng.random() < 0.7 else None
            qual = [(7 * (j // 11) + 3 * i) % 41 + 1 for j in range(len(seq))]
            recs.append([f"{ctg}_r{i}", st, flag, rng.choice([0, 3, 20, 60, 60, 60]), "".join(f"{l}{o}" for l, o in cig), seq, qual, hp])
        # supplementary alignments that share the query name with an earlier nearby read (rows are keyed by name)
        for _ in range(3):
            i = rng.randrange(len(recs) // 4, len(recs) // 2)
            j = min(len(recs) - 1, i + rng.randrange(1, 6))
            recs[j][0] = recs[i][0]
            recs[j][2] = (recs[j][2] & 16) | 2048
        reads[ctg] = recs
    vcf.sort(key=lambda r: (r[0], r[1]))
    return contigs, refs, reads, vcf


def vcf_text(vcf):
    out = ["##fileformat=VCFv4.2\n", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n"]
    for (ctg, pos, ref, alt, qual, gt) in vcf:
        out.append(f"{ctg}\t{pos}\t.\t{ref}\t{alt}\t{qual}\tPASS\t.\tGT:GQ\t{gt}:{int(qual)}\n")
    return "".join(out)


# ------------------------------------------------------------------ stubs
class _Recorder:
    files = []

    class _Node:
        def __init__(self):
            self.chunks = []

        def append(self, a):
            self.chunks.append(np.array(a))

    class _File:
        def __init__(self, path):
            self.path = path
            self.root = types.SimpleNamespace()

        def create_earray(self, where, name, atom=None, shape=None, filters=None):
            setattr(self.root, name, _Recorder._Node())

        def close(self):
            _Recorder.files.append((os.path.basename(self.path), {k: np.concatenate(v.chunks) for k, v in vars(self.root).items()}))


def install_stubs():
    ps = types.ModuleType("pysam")
    ps.AlignmentFile = pysam_emul.AlignmentFile.open
    sys.modules["pysam"] = ps
    tb = types.ModuleType("tables")
    tb.Filters = lambda **kw: None
    tb.Atom = types.SimpleNamespace(from_dtype=lambda d: None)
    tb.StringAtom = lambda itemsize: None
    tb.open_file = lambda path, mode="w": _Recorder._File(path)
    sys.modules["tables"] = tb


def run_reference(contigs, reads, vcf, threads, max_coverage, tmp):
    install_stubs()
    sys.path.insert(0, REF)
    import make_predict_bins as mpb                     # the reference module, unmodified
    import select_hetesnp_homosnp as sel
    bam_dir = os.path.join(tmp, "bams"); os.makedirs(bam_dir, exist_ok=True)
    for ctg, _ in contigs:
        path = os.path.join(bam_dir, ctg + ".bam")
        open(path, "w").close()
        segs = [pysam_emul.Segment(*r) for r in reads[ctg]]
        pysam_emul.AlignmentFile.registry[bam_dir + "/" + ctg + ".bam"] = pysam_emul.AlignmentFile(ctg, segs)
    vpath = os.path.join(tmp, "pileup.vcf")
    with open(vpath, "w") as f:
        f.write(vcf_text(vcf))
    groups = sel.select_snp_multiprocess(vcf_file=vpath, quality_threshold=19, adjacent_size=5, support_quality=14, nthreads=threads)
    gtab = {c: np.array([[int(it.position) for it in g] for g in gs], np.int64).reshape(-1, 11) for c, gs in groups.items()}
    out = os.path.join(tmp, f"out_{threads}_{max_coverage}"); os.makedirs(out, exist_ok=True)
    args = types.SimpleNamespace(pileup_vcf=vpath, low_quality_threshold=19, adjacent_size=5, pileup_flanking_size=16, hete_support_quality=14,
                                 threads=threads, bams=bam_dir, max_coverage=max_coverage, max_pileup_depth=None, max_haplotype_depth=None,
                                 output=out)
    _Recorder.files = []
    mpb.Run(args)
    return gtab, list(_Recorder.files)


def main():
    contigs, refs, reads, vcf = make_input()
    save = {"vcf_text": np.array(vcf_text(vcf)), "contigs": np.array([c for c, _ in contigs]), "contig_lens": np.array([l for _, l in contigs])}
    for ctg, _ in contigs:
        recs = reads[ctg]
        save[f"reads_{ctg}_name"] = np.array([r[0] for r in recs])
        save[f"reads_{ctg}_pos"] = np.array([r[1] for r in recs], np.int32)
        save[f"reads_{ctg}_flag"] = np.array([r[2] for r in recs], np.uint16)
        save[f"reads_{ctg}_mapq"] = np.array([r[3] for r in recs], np.uint8)
        save[f"reads_{ctg}_cigar"] = np.array([r[4] for r in recs])
        save[f"reads_{ctg}_seq"] = np.array([r[5] for r in recs])
        save[f"reads_{ctg}_qual"] = np.concatenate([np.asarray(r[6], np.uint8) for r in recs])
        save[f"reads_{ctg}_hp"] = np.array([r[7] or 0 for r in recs], np.uint8)
        save[f"ref_{ctg}"] = np.frombuffer(refs[ctg].encode(), np.uint8)
    with tempfile.TemporaryDirectory() as tmp:
        for tag, threads, maxcov in (("t3c150", 3, 150), ("t2c34", 2, 34)):
            gtab, files = run_reference(contigs, reads, vcf, threads, maxcov, tmp)
            save[f"{tag}_group_contigs"] = np.array(sorted(gtab))
            for c, t in gtab.items():
                save[f"{tag}_groups_{c}"] = t
            save[f"{tag}_files"] = np.array([f for f, _ in files])
            for fname, arrs in files:
                for k, v in arrs.items():
                    save[f"{tag}_file_{fname}_{k}"] = v if v.dtype.kind in "US" else v.astype(np.int32)
            print(tag, {c: t.shape for c, t in gtab.items()}, [(f, a["pileup_sequences"].shape) for f, a in files])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hapgroups_small.npz"), **save)
    print("wrote hapgroups_small.npz", os.path.getsize(os.path.join(ROOT, "tests", "golden", "hapgroups_small.npz")))


if __name__ == "__main__":
    main()
