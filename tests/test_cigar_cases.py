"""CIGARs outside SURVEY appendix B.4 and the htslib depth cap (appendix B.3).

CPU: the oracle's mpileup restatement reproduces the hand-worked rows of tests/golden/cigar_cases.txt; the depth-cap
restatement drops exactly the reads the published push rule refuses.  GPU: counts / flags / windows from the same reads are
bit-exact against the oracle chain, for the hand-worked cases, for random non-B.4 CIGARs and for pile-ups deeper than the cap.
"""
import numpy as np
import pytest

from conftest import GOLDEN, oracle_s1


def load_cases():
    reads, rows = [], []
    for line in (GOLDEN / "cigar_cases.txt").read_text().splitlines():
        if line.startswith("@"):
            _, pos, flag, mapq, cig, seq = line.split()
            reads.append((int(pos), int(flag), int(mapq), cig, seq))
        elif line.startswith(">"):
            _, pos1, depth, bases = line.split()
            rows.append(f"ctg1\t{pos1}\tN\t{depth}\t{bases}\t{'~' * int(depth)}")
    return reads, rows


def test_oracle_reproduces_hand_worked_rows(orc, tmp_path):
    from nanosnp_b200.reads import from_records
    recs, rows = load_cases()
    path = str(tmp_path / "c.mpileup")
    n, _ = orc.mpileup_text(from_records(recs), "ctg1", path)
    got = open(path).read().splitlines()
    assert got == rows, "\n".join(f"{a!r} != {b!r}" for a, b in zip(got, rows) if a != b)
    assert n == len(rows)


def test_canonicalize_merges_adjacent_runs():
    from nanosnp_b200.reads import canonicalize_cigars, from_records
    rd = canonicalize_cigars(from_records([(30, 0, 60, "2M1D2D2M1I2I2M", "ACGTAAACC"), (40, 0, 60, "3M", "ACG"), (41, 0, 60, "1M1M2D1D", "AC")]))
    def cig(i):
        return [(int(c) >> 4, "MIDNSHP=X"[int(c) & 15]) for c in rd.cigar[rd.cigar_off[i]:rd.cigar_off[i + 1]]]
    assert cig(0) == [(2, "M"), (3, "D"), (2, "M"), (3, "I"), (2, "M")]
    assert cig(1) == [(3, "M")] and cig(2) == [(2, "M"), (3, "D")]


def deep_reads(n_same=200, n_more=120, seed=3):
    """Pile-up deeper than the cap: n_same reads starting at one position, then reads at following positions."""
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n_same):
        L = int(rng.integers(40, 400))
        recs.append((100, 16 * int(rng.integers(0, 2)), 60, f"{L}M", "".join("ACGT"[b] for b in rng.integers(0, 4, L))))
    for i in range(n_more):
        L = int(rng.integers(40, 400))
        p = 100 + 1 + i // 3                                  # three reads per start: the first of each start is exempt
        recs.append((p, 16 * int(rng.integers(0, 2)), 60 if i % 7 else 5, f"{L}M", "".join("ACGT"[b] for b in rng.integers(0, 4, L))))
    return recs


def test_depth_cap_restatement_rule(orc):
    """B.3 push rule by brute force: dropped iff an earlier passing read starts at the same position AND the pool
    (pushed reads with end >= pos, + 2 bookkeeping nodes) already exceeds max_depth."""
    from nanosnp_b200.reads import from_records
    recs = deep_reads()
    rd = from_records(recs)
    got = orc.depth_cap(rd, max_depth=144)
    pushed, exp, last = [], [], None
    for (p, f, q, cg, s) in recs:
        if q < 20:
            exp.append(0); continue
        alive = sum(1 for e in pushed if e >= p)
        drop = last == p and alive + 2 > 144
        exp.append(int(drop))
        if not drop:
            pushed.append(p + len(s)); last = p
    assert list(got) == exp
    assert 50 < sum(exp) < len(recs) - 143          # 143 reads fill the pool; later same-start reads are refused
    assert orc.depth_cap(rd, max_depth=0).sum() == 0
    # deeper columns than the cap can exist (first read of each start is always pushed): the cap is not a truncation
    assert orc.depth_cap(from_records([(i, 0, 60, "500M", "A" * 500) for i in range(300)]), max_depth=144).sum() == 0


# ------------------------------------------------------------------------------------------------ GPU
def random_cigar_reads(seed, n_reads=600, L=6000):
    """Reads with CIGARs far outside B.4: leading / trailing / adjacent I-D, N skips, clips, =/X, reads made of indels."""
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n_reads):
        pos = int(rng.integers(0, L - 700))
        ops, seq, ref_left, prev = [], [], 650, None
        if rng.random() < 0.3: ops.append((int(rng.integers(1, 30)), "S"))
        if rng.random() < 0.2: ops.insert(0, (int(rng.integers(1, 9)), "H"))
        n_ops = int(rng.integers(1, 40))
        for k in range(n_ops):
            choices = [c for c in "MMM=XIDDIN" if c != prev or c in "M"]      # no adjacent equal indel / skip ops (canonical input)
            c = choices[int(rng.integers(0, len(choices)))]
            ln = int(rng.integers(1, 70 if c in "ID" else 25 if c != "N" else 12))
            if c in "MDN=X":
                if ref_left - ln < 1: break
                ref_left -= ln
            if prev == c: continue
            ops.append((ln, c)); prev = c
        if not any(c in "M=XDN" for _, c in ops): ops.append((5, "M"))
        if rng.random() < 0.3: ops.append((int(rng.integers(1, 30)), "S"))
        merged = []
        for ln, c in ops:
            if merged and merged[-1][1] == c: merged[-1] = (merged[-1][0] + ln, c)
            else: merged.append((ln, c))
        cig = "".join(f"{ln}{c}" for ln, c in merged)
        qlen = sum(ln for ln, c in merged if c in "MIS=X")
        s = "".join("ACGTN"[b] for b in rng.choice(5, qlen, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
        recs.append((pos, 16 * int(rng.integers(0, 2)), int(rng.choice([60, 60, 60, 3])), cig, s))
    recs.sort(key=lambda r: r[0])
    return recs


@pytest.fixture(scope="module")
def engine():
    import torch
    from nanosnp_b200.pipeline import PileupEngine
    assert torch.cuda.is_available()
    return PileupEngine("cuda:0")


def _gpu(engine, reads, ref):
    import torch
    rd = reads.to_torch(engine.device)
    rf = torch.from_numpy(np.ascontiguousarray(ref)).to(engine.device)
    pos, refbase, x, counts, flags = engine.candidate_windows(rd, rf)
    torch.cuda.synchronize()
    return pos.cpu().numpy(), x.cpu().numpy(), counts.cpu().numpy(), flags.cpu().numpy()


def _same_as_oracle(res, pos, x, counts, flags):
    cov = (res.flags & 1).astype(bool)
    assert np.array_equal(flags & 1, res.flags & 1)
    bad = np.nonzero((counts[cov] != res.counts[cov]).any(1))[0]
    assert bad.size == 0, (np.nonzero(cov)[0][bad[:5]], counts[cov][bad[:2]], res.counts[cov][bad[:2]])
    assert np.array_equal(flags, res.flags)
    assert np.array_equal(pos + 1, res.positions) and np.array_equal(x, res.windows)


@pytest.mark.gpu
def test_gpu_hand_worked_and_random_cigars(engine, orc, tmp_path):
    from nanosnp_b200 import _lib
    from nanosnp_b200.reads import canonicalize_cigars, from_records
    rng = np.random.default_rng(11)
    ref = np.frombuffer(bytes(rng.choice(list(b"ACGT"), 6000).astype(np.uint8)), np.uint8)
    recs, _ = load_cases()
    nopad = [r for r in recs if "P" not in r[3]]
    rd = canonicalize_cigars(from_records(nopad))
    _same_as_oracle(oracle_s1(orc, rd, ref[:100], tmp_path), *_gpu(engine, rd, ref[:100]))
    # unmerged runs and pads are refused, not silently counted differently
    for bad in ([(30, 0, 60, "2M1D2D2M", "ACGT")], [(50, 0, 60, "2M1P1I1P2M", "ACAGT")]):
        with pytest.raises(_lib.NsnpError) as e:
            _gpu(engine, from_records(bad), ref[:100])
        assert e.value.code == _lib.E_UNSUPPORTED
    _gpu(engine, rd, ref[:100])                                  # the status word was cleared: the engine keeps working
    for seed in (1, 2, 3):
        rd = from_records(random_cigar_reads(seed))
        _same_as_oracle(oracle_s1(orc, rd, ref, tmp_path), *_gpu(engine, rd, ref))


@pytest.mark.gpu
def test_gpu_depth_cap(engine, orc, tmp_path):
    """Pile-ups deeper than --max-depth: the GPU drops exactly the reads the restated push rule drops."""
    from nanosnp_b200.reads import from_records
    from nanosnp_b200.synth import SynthConfig, generate_host
    rng = np.random.default_rng(5)
    ref = np.frombuffer(bytes(rng.choice(list(b"ACGT"), 1200).astype(np.uint8)), np.uint8)
    rd = from_records(deep_reads())
    assert orc.depth_cap(rd).sum() > 50
    _same_as_oracle(oracle_s1(orc, rd, ref, tmp_path), *_gpu(engine, rd, ref))
    # 400x of short synthetic reads with indels: many starts share a position once the pool is full
    cfg = SynthConfig(contig_len=20_000, coverage=400.0, len_median=500, len_min=100, len_sigma=0.5, seed_reads=77)
    ref2, rd2 = generate_host(cfg)
    n_drop = int(orc.depth_cap(rd2).sum())
    assert n_drop > 100
    _same_as_oracle(oracle_s1(orc, rd2, ref2, tmp_path), *_gpu(engine, rd2, ref2))
    # with the cap disabled both sides count every read
    import ctypes as C
    from nanosnp_b200 import _lib
    from nanosnp_b200.pipeline import PileupEngine
    p = _lib.default_params(); p.max_depth = 0
    eng0 = PileupEngine("cuda:0", p)
    mp = str(tmp_path / "nocap.mpileup")
    orc.mpileup_text(rd2, "ctg1", mp, max_depth=0)
    _same_as_oracle(orc.s1_restate(mp, "ctg1", ref2), *_gpu(eng0, rd2, ref2))
