"""GPU parity, stage s1: pileup counts / flags / candidates / windows must be BIT-EXACT against the oracle chain
(mpileup restatement -> s1 restatement, the latter pinned to the reference's own binaries) and against the
committed outputs of the reference binaries."""
import numpy as np
import pytest

from conftest import oracle_s1

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import torch
    from nanosnp_b200.pipeline import PileupEngine
    assert torch.cuda.is_available()
    return PileupEngine("cuda:0")


def _run_gpu(engine, reads, ref, **kw):
    import torch
    rd = reads.to_torch(engine.device)
    rf = torch.from_numpy(np.ascontiguousarray(ref)).to(engine.device)
    pos, refbase, x, counts, flags = engine.candidate_windows(rd, rf, **kw)
    torch.cuda.synchronize()
    return pos.cpu().numpy(), refbase.cpu().numpy(), x.cpu().numpy(), counts.cpu().numpy(), flags.cpu().numpy()


def _compare(res, pos, x, counts, flags, lo=0, hi=None):
    hi = len(res.flags) if hi is None else hi
    exp_flags = res.flags[lo:hi]
    cov = (exp_flags & 1).astype(bool)
    bad = np.nonzero((flags & 1) != (exp_flags & 1))[0]
    assert bad.size == 0, f"covered differs at {bad[:10] + lo}"
    d = np.nonzero((counts[cov] != res.counts[lo:hi][cov]).any(1))[0]
    if d.size:
        p = np.nonzero(cov)[0][d[0]]
        raise AssertionError(f"counts differ at 0-based {p + lo}: gpu {counts[p]} oracle {res.counts[lo + p]} ({d.size} rows)")
    assert (counts[~cov] == 0).all()
    bad = np.nonzero(flags != exp_flags)[0]
    assert bad.size == 0, f"gate differs at {bad[:10] + lo}"


def test_small_case_bit_exact(engine, orc, small_case, tmp_path):
    res = oracle_s1(orc, small_case["reads"], small_case["ref"], tmp_path)
    pos, refbase, x, counts, flags = _run_gpu(engine, small_case["reads"], small_case["ref"])
    _compare(res, pos, x, counts, flags)
    # against what the reference's own binaries produced (tests/golden/s1_small.npz)
    assert np.array_equal(pos + 1, small_case["site_pos"])
    assert np.array_equal(x, small_case["windows"])
    assert np.array_equal(refbase, small_case["site_refbase"])


CASES = {
    "plain": dict(contig_len=120_000, coverage=30.0),
    "eqx_nbases": dict(contig_len=60_000, coverage=25.0, use_eqx=True, nbase_rate=0.01),
    "gaps_refN_lower": dict(contig_len=80_000, coverage=20.0, gap_period=9000, gap_len=200, ref_n_period=7000, ref_n_len=40,
                            ref_lower_period=3000, ref_lower_len=500, nbase_rate=0.003),
    "deep_short_reads": dict(contig_len=30_000, coverage=100.0, len_median=600, len_min=100, len_sigma=0.4),
    "low_cov": dict(contig_len=100_000, coverage=8.0, snp_rate=5e-3),
    "long_indels": dict(contig_len=50_000, coverage=30.0, long_indel_rate=0.05, indel_mean=6.0),
    "tiny": dict(contig_len=700, coverage=12.0, len_median=200, len_min=50),
}


@pytest.mark.parametrize("name", list(CASES))
def test_synthetic_cases_bit_exact(engine, orc, tmp_path, name):
    from nanosnp_b200.synth import SynthConfig, generate_host
    cfg = SynthConfig(seed_ref=7, seed_var=8, seed_reads=9, **CASES[name])
    ref, reads = generate_host(cfg)
    res = oracle_s1(orc, reads, ref, tmp_path)
    pos, refbase, x, counts, flags = _run_gpu(engine, reads, ref)
    _compare(res, pos, x, counts, flags)
    assert np.array_equal(pos + 1, res.positions), (len(pos), len(res.positions))
    assert np.array_equal(x, res.windows)
    assert np.array_equal(refbase, np.char.upper(ref[pos].view("S1")).view(np.uint8))


def test_region_with_halo_and_standalone_select(engine, orc, tmp_path):
    """A shard [s,e) computed on [s-16, e+16) emits exactly the oracle's sites with s <= c < e;
    nsnp_select_candidates with recompute_gate reproduces the gate from the count rows alone."""
    import torch
    from nanosnp_b200.synth import SynthConfig, generate_host
    cfg = SynthConfig(contig_len=90_000, coverage=22.0, seed_ref=21, seed_var=22, seed_reads=23, len_median=3000, len_min=200)
    ref, reads = generate_host(cfg)
    res = oracle_s1(orc, reads, ref, tmp_path)
    got = []
    for s, e in ((0, 30_011), (30_011, 61_000), (61_000, 90_000)):
        rs, re_ = max(0, s - 16), min(cfg.contig_len, e + 16)
        pos, refbase, x, counts, flags = _run_gpu(engine, reads, ref, region_start=rs, region_len=re_ - rs, emit_start=s, emit_end=e)
        _compare(res, pos, x, counts, flags, rs, re_)
        got.append(pos)
        sel = (res.positions - 1 >= s) & (res.positions - 1 < e)
        assert np.array_equal(x, res.windows[sel])
    assert np.array_equal(np.concatenate(got) + 1, res.positions)
    # stand-alone select: wipe the GATE bit, recompute from counts
    rd = reads.to_torch(engine.device); rf = torch.from_numpy(ref).to(engine.device)
    counts, flags = engine.pileup_counts(rd, rf)
    flags2 = flags & 1
    pos, n_dev = engine.select(flags2, rf, 0, 0, cfg.contig_len, cfg.contig_len, counts=counts, recompute_gate=True)
    n = int(n_dev.item())
    assert np.array_equal(pos[:n].cpu().numpy() + 1, res.positions)
    assert torch.equal(flags2, flags)


def test_empty_and_filtered_inputs(engine):
    import torch
    from nanosnp_b200.reads import from_records
    ref = np.frombuffer(b"ACGT" * 50, np.uint8)
    # every read fails a filter: unmapped, secondary, supplementary, low MAPQ
    recs = [(10, 4, 60, "50M", "A" * 50), (12, 256, 60, "50M", "C" * 50), (14, 2048, 60, "50M", "G" * 50), (16, 0, 19, "50M", "T" * 50)]
    pos, refbase, x, counts, flags = _run_gpu(engine, from_records(recs), ref)
    assert len(pos) == 0 and (counts == 0).all() and (flags == 0).all()
    # QCFAIL / DUP are kept (excl-flags 2316, not samtools' default 1796)
    recs = [(0, 512, 60, "200M", "ACGT" * 50), (0, 1024 + 16, 60, "200M", "ACGT" * 50)]
    pos, refbase, x, counts, flags = _run_gpu(engine, from_records(recs), ref)
    assert (flags & 1).all() and (counts[:, [0, 1, 2, 3]].sum(1) == -1).all() and (counts[:, [9, 10, 11, 12]].sum(1) == -1).all()


def test_device_generator_matches_host(engine):
    import torch
    from nanosnp_b200.synth import SynthConfig, generate_host, generate_device
    cfg = SynthConfig(contig_len=200_000, coverage=15.0, nbase_rate=0.004, gap_period=50_000, gap_len=300, use_eqx=True)
    ref_h, rd_h = generate_host(cfg)
    ref_d, rd_d = generate_device(cfg, engine.device)
    assert np.array_equal(ref_h, ref_d.cpu().numpy())
    rd_d = rd_d.to_numpy()
    for f in ("pos", "flag", "mapq", "cigar_off", "cigar", "seq_off", "seq2", "nmask"):
        assert np.array_equal(getattr(rd_h, f), getattr(rd_d, f)), f


def test_cigar16_words_give_identical_counts(engine):
    """struct nsnp_reads.cigar_bits = 16: the same CIGAR values shipped as uint16 (host numpy and device torch packing)."""
    import torch
    from nanosnp_b200.reads import cigar16
    from nanosnp_b200.synth import SynthConfig, generate_host
    cfg = SynthConfig(contig_len=150_000, coverage=25.0, seed_ref=3, seed_var=4, seed_reads=5, nbase_rate=0.003)
    ref, reads = generate_host(cfg)
    want = _run_gpu(engine, reads, ref)
    r16 = cigar16(reads)
    assert r16.cigar.dtype == np.uint16 and r16.as_struct().cigar_bits == 16 and reads.as_struct().cigar_bits == 32
    got = _run_gpu(engine, r16, ref)
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    d16 = cigar16(reads.to_torch(engine.device))
    assert d16.cigar.dtype == torch.int16 and d16.as_struct().cigar_bits == 16
    rf = torch.from_numpy(np.ascontiguousarray(ref)).to(engine.device)
    pos, refbase, x, counts, flags = engine.candidate_windows(d16, rf)
    assert np.array_equal(counts.cpu().numpy(), want[3]) and np.array_equal(pos.cpu().numpy(), want[0])
    # a length >= 4096 keeps the 32-bit words
    from nanosnp_b200.reads import from_records
    big = from_records([(0, 0, 60, "5000M", "A" * 5000)])
    assert cigar16(big).cigar.dtype == np.uint32
