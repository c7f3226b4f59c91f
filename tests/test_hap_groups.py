"""HaplotypeModel s4 (BASELINE configs[4]): host ports, oracle restatement and the GPU read-matrix kernel against
tests/golden/hapgroups_small.npz, which the reference's own code produced (tests/golden/make_golden_hapgroups.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nanosnp_b200 import hap_groups as hg                      # noqa: E402
from nanosnp_b200.bam import ReadAux, qname_hash               # noqa: E402
from nanosnp_b200.reads import from_records                    # noqa: E402
from oracle import pysam_emul                                  # noqa: E402
from oracle.hap_groups_restate import canonical_rows, subgroup_matrices   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "hapgroups_small.npz")
RUNS = (("t3c150", 3, 150), ("t2c34", 2, 34))


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD, allow_pickle=False)


def records(z, ctg):
    qual = z[f"reads_{ctg}_qual"]
    out, o = [], 0
    for i in range(len(z[f"reads_{ctg}_pos"])):
        seq = str(z[f"reads_{ctg}_seq"][i])
        out.append((str(z[f"reads_{ctg}_name"][i]), int(z[f"reads_{ctg}_pos"][i]), int(z[f"reads_{ctg}_flag"][i]), int(z[f"reads_{ctg}_mapq"][i]),
                    str(z[f"reads_{ctg}_cigar"][i]), seq, qual[o:o + len(seq)], int(z[f"reads_{ctg}_hp"][i]) or None))
        o += len(seq)
    return out


def packed(recs):
    rd = from_records([(r[1], r[2], r[3], r[4], r[5]) for r in recs])
    qual = np.zeros(rd.n_bases + 16, np.uint8)
    for so, r in zip(rd.seq_off, recs):
        qual[so:so + len(r[5])] = r[6]
    aux = ReadAux(qual, np.array([r[7] or 0 for r in recs], np.uint8), np.array([qname_hash(r[0]) for r in recs], np.uint64), [r[0] for r in recs])
    return rd, aux


def vcf_file(z, tmp_path):
    p = tmp_path / "pileup.vcf"
    p.write_text(str(z["vcf_text"]))
    return str(p)


def assert_same_groups(got, want, centre):
    """4 x [n, depth, L] against 4 x [n, depth, L]: equal up to the order of rows inside one HP class."""
    assert got[0].shape == want[0].shape
    for g in range(got[0].shape[0]):
        a = canonical_rows(*[m[g] for m in got], centre)
        b = canonical_rows(*[m[g] for m in want], centre)
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("tag,threads,maxcov", RUNS)
def test_group_selection_matches_reference(gold, tmp_path, tag, threads, maxcov):
    got = hg.select_snp_multiprocess(vcf_file(gold, tmp_path), 19, 5, 14, nthreads=threads)
    want = {str(c): gold[f"{tag}_groups_{c}"] for c in gold[f"{tag}_group_contigs"]}
    assert sorted(got) == sorted(want)                       # incl. the contigs the reference's slice bug loses
    for c in want:
        np.testing.assert_array_equal(got[c], want[c])
    everything = hg.select_snp_multiprocess(vcf_file(gold, tmp_path), 19, 5, 14, nthreads=threads, keep_reference_bug=False)
    assert set(everything) >= set(got) and len(everything) == 3


def test_plan_subgroups_rules():
    g = np.arange(250)[:, None] * 50 + np.arange(11)[None, :]            # dense: cut every 100 groups
    assert hg.plan_subgroups(g) == [(0, 100), (100, 200), (200, 250)]
    g2 = g.copy(); g2[30:] += 5000; g2[31:] += 5000                      # two gaps > 1000
    assert hg.plan_subgroups(g2) == [(0, 30), (30, 31), (31, 131), (131, 231), (231, 250)]
    assert hg.plan_subgroups(g[:1]) == [(0, 1)] and hg.plan_subgroups(g[:0]) == []
    assert hg.plan_chunks(10, 3) == [(0, 4), (4, 8), (8, 10)] and hg.plan_chunks(0, 3) == []


def oracle_files(z, tag, threads, maxcov, tmp_path):
    """The reference's Run() restated with the oracle: {file name: arrays}."""
    groups = hg.select_snp_multiprocess(vcf_file(z, tmp_path), 19, 5, 14, nthreads=threads)
    files = {}
    for ctg, g in groups.items():
        sam = pysam_emul.AlignmentFile(ctg, [pysam_emul.Segment(*r) for r in records(z, ctg)])
        for lo, hi in hg.plan_chunks(len(g), threads):
            res = []
            for a, b in hg.plan_subgroups(g[lo:hi]):
                res += subgroup_matrices(sam, ctg, g[lo + a:lo + b], maxcov, 16)
            if not res:
                continue
            res.sort(key=lambda r: r["positions"][5])
            d = max(r["hap"][0].shape[0] for r in res)
            pad = lambda m: np.pad(m, ((0, d - m.shape[0]), (0, 0)), constant_values=-2)
            arrs = {}
            for k, name in enumerate(hg.NAMES):
                arrs["haplotype_" + name] = np.stack([pad(r["hap"][k]) for r in res]).astype(np.int32)
                arrs["pileup_" + name] = np.stack([pad(r["pile"][k]) for r in res]).astype(np.int32)
            arrs["candidate_positions"] = np.array([[f"{ctg}:{r['positions'][5]}"] for r in res])
            arrs["haplotype_positions"] = np.array([[f"{ctg}:{p}" for p in r["positions"]] for r in res])
            files[f"{ctg}_{res[0]['positions'][5]}_{res[-1]['positions'][5]}.bin"] = arrs
    return files


def check_files(z, tag, files):
    want_names = sorted(str(f) for f in z[f"{tag}_files"])
    assert sorted(files) == want_names
    for f in want_names:
        w = {k: z[f"{tag}_file_{f}_{k}"] for k in files[f]}
        np.testing.assert_array_equal(files[f]["candidate_positions"], w["candidate_positions"])
        np.testing.assert_array_equal(files[f]["haplotype_positions"], w["haplotype_positions"])
        assert_same_groups([files[f]["haplotype_" + n] for n in hg.NAMES], [w["haplotype_" + n] for n in hg.NAMES], 5)
        assert_same_groups([files[f]["pileup_" + n] for n in hg.NAMES], [w["pileup_" + n] for n in hg.NAMES], 16)


@pytest.mark.parametrize("tag,threads,maxcov", RUNS)
def test_oracle_restatement_matches_reference(gold, tmp_path, tag, threads, maxcov):
    check_files(gold, tag, oracle_files(gold, tag, threads, maxcov, tmp_path))


def test_golden_exercises_the_quirks(gold):
    """The fixture must contain what the kernel has to reproduce: dropped sub-groups, deletions, skips, shared query names."""
    f0 = str(gold["t3c150_files"][0])
    seq = np.concatenate([gold[f"t3c150_file_{f}_pileup_sequences"].ravel() for f in gold["t3c150_files"]])
    assert (seq == -1).any() and (seq == -2).any() and (seq == 0).any()
    n_groups = sum(len(gold[f"t3c150_groups_{c}"]) for c in gold["t3c150_group_contigs"])
    n_out = sum(gold[f"t3c150_file_{f}_candidate_positions"].shape[0] for f in gold["t3c150_files"])
    assert 0 < n_out < n_groups
    names = gold["reads_chrA_name"]
    assert len(set(names.tolist())) < len(names)
    assert f0.endswith(".bin")


# ------------------------------------------------------------------------------------------------ GPU
def gpu_files(z, threads, maxcov, tmp_path, via_bam):
    import torch
    from nanosnp_b200.bam import write_bam
    groups = hg.select_snp_multiprocess(vcf_file(z, tmp_path), 19, 5, 14, nthreads=threads)
    files = {}
    lens = dict(zip([str(c) for c in z["contigs"]], [int(l) for l in z["contig_lens"]]))
    for ctg, g in groups.items():
        rd, aux = packed(records(z, ctg))
        if via_bam:
            bdir = tmp_path / "bams"; bdir.mkdir(exist_ok=True)
            write_bam(str(bdir / f"{ctg}.bam"), [(ctg, lens[ctg])], {ctg: rd}, aux={ctg: aux})
            al = hg.load_contig(str(bdir / f"{ctg}.bam"), ctg)
        else:
            al = hg.upload_alignments(rd, aux)
        chunks = hg.plan_chunks(len(g), threads)
        subs = [(lo + a, lo + b) for lo, hi in chunks for a, b in hg.plan_subgroups(g[lo:hi])]
        gm = hg.group_matrices(al, g, subs, maxcov, 16)
        for lo, hi in chunks:
            arrs = hg.chunk_arrays(gm, ctg, lo, hi)
            if arrs is not None:
                first = arrs["candidate_positions"][0, 0].split(":")[1]; last = arrs["candidate_positions"][-1, 0].split(":")[1]
                files[f"{ctg}_{first}_{last}.bin"] = arrs
    torch.cuda.synchronize()
    return files


@pytest.mark.gpu
@pytest.mark.parametrize("via_bam", [False, True])
@pytest.mark.parametrize("tag,threads,maxcov", RUNS)
def test_gpu_matrices_match_reference(gold, tmp_path, tag, threads, maxcov, via_bam):
    check_files(gold, tag, gpu_files(gold, threads, maxcov, tmp_path, via_bam))


@pytest.mark.gpu
def test_gpu_run_writes_the_files_the_dataset_reads(gold, tmp_path):
    """make_predict_bins CLI equivalent end to end: BAM directory + VCF -> .npz files -> nanosnp_b200.haplotype.TestDataset."""
    from nanosnp_b200.bam import write_bam
    from nanosnp_b200.haplotype import TestDataset
    lens = dict(zip([str(c) for c in gold["contigs"]], [int(l) for l in gold["contig_lens"]]))
    bdir = tmp_path / "bams"; bdir.mkdir()
    refs = {}
    for ctg in lens:
        rd, aux = packed(records(gold, ctg))
        write_bam(str(bdir / f"{ctg}.bam"), [(ctg, lens[ctg])], {ctg: rd}, aux={ctg: aux})
        refs[ctg] = gold[f"ref_{ctg}"]
    out = tmp_path / "bins"
    written = hg.run(vcf_file(gold, tmp_path), str(bdir), str(out), pileup_flanking_size=16, threads=3, max_pileup_depth=90, max_haplotype_depth=90)
    assert sorted(os.path.basename(w)[:-4] + ".bin" for w in written) == sorted(str(f) for f in gold["t3c150_files"])
    ds = TestDataset(written[0], refs)
    pos, xp, xh = ds[0]
    assert xp.shape == (105, 33) and xh.shape == (105, 11) and ":" in pos


@pytest.mark.gpu
def test_gpu_fused_s4_s5_equals_the_file_path(gold, tmp_path):
    """predict_from_bams (VCF + BAMs -> rows, nothing on disk in between) == hap_groups.run -> .npz files -> haplotype.predict."""
    from nanosnp_b200 import haplotype as G
    from nanosnp_b200.bam import write_bam
    from oracle.hap_restate import HaplotypeModelOracle
    lens = dict(zip([str(c) for c in gold["contigs"]], [int(l) for l in gold["contig_lens"]]))
    bdir = tmp_path / "bams"; bdir.mkdir()
    fa = tmp_path / "ref.fa"
    with open(fa, "wb") as f:
        for ctg in lens:
            rd, aux = packed(records(gold, ctg))
            write_bam(str(bdir / f"{ctg}.bam"), [(ctg, lens[ctg])], {ctg: rd}, aux={ctg: aux})
            f.write(b">" + ctg.encode() + b"\n" + bytes(gold[f"ref_{ctg}"]) + b"\n")
    net = G.LSTMNetwork().to("cuda")
    net.load_state_dict(HaplotypeModelOracle(seed=5).state_dict())
    vcf = vcf_file(gold, tmp_path)
    out = tmp_path / "bins"
    hg.run(vcf, str(bdir), str(out), pileup_flanking_size=16, threads=3, max_pileup_depth=24, max_haplotype_depth=24)
    n1 = G.predict(net, str(out), str(fa), 7, 33, 11, str(tmp_path / "a.csv"), "cuda")
    n2 = G.predict_from_bams(net, vcf, str(bdir), str(fa), str(tmp_path / "b.csv"), batch_size=7, max_depth=24, threads=3)
    a = sorted((tmp_path / "a.csv").read_text().splitlines()); b = sorted((tmp_path / "b.csv").read_text().splitlines())
    assert n1 == n2 > 0 and a == b


def test_bam_reader_keeps_qualities_hp_tags_and_name_hashes(gold, tmp_path):
    """nsnp_bam_keep_aux / nsnp_bam_take_aux (host C++): what the s4 stage needs beside the packed reads survives a BAM round trip."""
    from nanosnp_b200.bam import BamReader, write_bam
    rd, aux = packed(records(gold, "chrB"))
    path = str(tmp_path / "chrB.bam")
    write_bam(path, [("chrB", 14000)], {"chrB": rd}, aux={"chrB": aux})
    with BamReader(path, keep_aux=True) as r:
        (_, name, got), = list(r.contigs())
        ax = r.aux
    assert name == "chrB" and got.n_reads == rd.n_reads
    np.testing.assert_array_equal(got.pos, rd.pos); np.testing.assert_array_equal(got.seq_off, rd.seq_off)
    np.testing.assert_array_equal(ax.hp, aux.hp); np.testing.assert_array_equal(ax.qhash, aux.qhash)
    for i in range(rd.n_reads):
        so = int(rd.seq_off[i]); n = len(str(gold["reads_chrB_seq"][i]))
        np.testing.assert_array_equal(ax.qual[so:so + n], aux.qual[so:so + n])
    with BamReader(path) as r:                                      # without keep_aux nothing is kept and take_aux refuses
        list(r.contigs())
        assert r.aux is None


@pytest.mark.gpu
@pytest.mark.parametrize("seed,threads,maxcov,adj,flank", [(23, 3, 150, 5, 16), (47, 1, 30, 5, 16), (5, 4, 26, 5, 16), (9, 2, 150, 3, 5), (13, 3, 40, 7, 32)])
def test_gpu_matrices_match_oracle_on_fresh_inputs(tmp_path, seed, threads, maxcov, adj, flank):
    """Other seeds than the golden file's: the GPU path against the (golden-pinned) oracle restatement on the same input."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_hapgroups as M
    contigs, refs, reads, vcf = M.make_input(seed)
    vpath = tmp_path / "p.vcf"; vpath.write_text(M.vcf_text(vcf))
    groups = hg.select_snp_multiprocess(str(vpath), 19, adj, 14, nthreads=threads)
    assert groups
    n_checked = 0
    c_h = adj
    for ctg, g in groups.items():
        recs = [tuple(r) for r in reads[ctg]]
        sam = pysam_emul.AlignmentFile(ctg, [pysam_emul.Segment(*r) for r in recs])
        rd, aux = packed(recs)
        al = hg.upload_alignments(rd, aux)
        chunks = hg.plan_chunks(len(g), threads)
        subs = [(lo + a, lo + b) for lo, hi in chunks for a, b in hg.plan_subgroups(g[lo:hi])]
        gm = hg.group_matrices(al, g, subs, maxcov, flank)
        want = []
        for a, b in subs:
            want += subgroup_matrices(sam, ctg, g[a:b], maxcov, flank)
        assert [list(p) for p in gm.positions] == [w["positions"] for w in want]
        hap = [h.cpu().numpy() for h in gm.hap]; pile = [p.cpu().numpy() for p in gm.pile]
        for i, w in enumerate(want):
            d = w["hap"][0].shape[0]
            assert gm.depth[i] == d
            if d == 0:
                continue
            for got, exp, centre in (([m[i, :d] for m in hap], w["hap"], c_h), ([m[i, :d] for m in pile], w["pile"], flank)):
                for x, y in zip(canonical_rows(*got, centre), canonical_rows(*exp, centre)):
                    np.testing.assert_array_equal(x, y)
            assert (hap[0][i, d:] == -2).all() and (pile[3][i, d:] == -2).all()
            n_checked += 1
    assert n_checked > 5
