"""BAM front end (host): BGZF + BAM round trip of synthetic reads, incl. N bases, odd lengths and the CG-tag long CIGAR."""
import numpy as np

from nanosnp_b200.bam import bgzf_compress, bgzf_decompress, read_bam, write_bam
from nanosnp_b200.reads import from_records
from nanosnp_b200.synth import SynthConfig, generate_host


def _same(a, b):
    n = a.n_reads
    assert n == b.n_reads
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.flag, b.flag) and np.array_equal(a.mapq, b.mapq)
    assert np.array_equal(a.cigar_off, b.cigar_off) and np.array_equal(a.cigar, b.cigar)
    for i in range(n):
        cg = a.cigar[a.cigar_off[i]:a.cigar_off[i + 1]]; ops = cg & 15
        l = int((cg >> 4)[(ops == 0) | (ops == 1) | (ops == 4) | (ops == 7) | (ops == 8)].sum())
        ka = int(a.seq_off[i]) + np.arange(l); kb = int(b.seq_off[i]) + np.arange(l)
        na = ((a.nmask[ka >> 3] >> (ka & 7)) & 1) if a.nmask is not None else np.zeros(l, np.uint8)
        nb = ((b.nmask[kb >> 3] >> (kb & 7)) & 1) if b.nmask is not None else np.zeros(l, np.uint8)
        assert np.array_equal(na, nb)
        ca = (a.seq2[ka >> 2] >> (2 * (ka & 3))) & 3; cb = (b.seq2[kb >> 2] >> (2 * (kb & 3))) & 3
        assert np.array_equal(ca[na == 0], cb[nb == 0])


def test_bgzf_round_trip(tmp_path):
    raw = np.random.default_rng(0).integers(0, 255, 300_000, dtype=np.uint8).tobytes()
    p = tmp_path / "x.bgzf"; p.write_bytes(bgzf_compress(raw))
    assert bgzf_decompress(str(p)) == raw


def test_bam_round_trip_synthetic(tmp_path):
    cfg = SynthConfig(contig_len=40_000, coverage=6.0, len_median=1500, len_min=100, nbase_rate=0.01)
    ref, rd = generate_host(cfg)
    cfg2 = SynthConfig(contig_len=9_000, coverage=4.0, len_median=800, len_min=100, seed_reads=5, contig="ctg2")
    _, rd2 = generate_host(cfg2)
    path = str(tmp_path / "t.bam")
    write_bam(path, [("ctg1", 40_000), ("empty", 500), ("ctg2", 9_000)], {"ctg1": rd, "ctg2": rd2})
    refs, got = read_bam(path)
    assert refs == [("ctg1", 40_000), ("empty", 500), ("ctg2", 9_000)] and set(got) == {"ctg1", "ctg2"}
    _same(rd, got["ctg1"]); _same(rd2, got["ctg2"])
    assert (got["ctg1"].seq_off % 16 == 0).all()
    _, only = read_bam(path, contigs={"ctg2"})
    assert set(only) == {"ctg2"}


def test_long_cigar_in_cg_tag(tmp_path):
    recs = [(5, 0, 60, "3S10M2I5M1D7M", "ACGTNACGTACGTACGTACGTACGTAC"), (9, 16, 30, "21M", "ACGTACGTACGTACGTACGTA")]
    rd = from_records(recs)
    path = str(tmp_path / "cg.bam")
    write_bam(path, [("c", 100)], {"c": rd}, long_cigar_as_tag=4)      # first read has 6 ops -> stored in CG:B,I
    _, got = read_bam(path)
    _same(rd, got["c"])
