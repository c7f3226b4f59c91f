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


# ---------------------------------------------------------------------------------------------------------------
# Byte-level fixture built here from the SAM/BAM specification (sections 4.1, 4.2), NOT through write_bam.
def _bgzf_block(payload: bytes, extra_first: bytes = b"") -> bytes:
    import struct, zlib
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(payload) + co.flush()
    xlen = 6 + len(extra_first)
    bsize = 12 + xlen + len(comp) + 8 - 1
    return (struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, xlen) + extra_first + struct.pack("<BBHH", 66, 67, 2, bsize) + comp
            + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload)))


def _bam_record(ref_id, pos, mapq, flag, cigar, seq, tags=b"", name=b"q\0"):
    import struct
    codes = "=ACMGRSVTWYHKDBN"
    ops = "MIDNSHP=X"
    cig = b"".join(struct.pack("<I", (n << 4) | ops.index(o)) for n, o in cigar)
    nib = [codes.index(c) for c in seq] + ([0] if len(seq) & 1 else [])
    sq = bytes((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))
    body = struct.pack("<iiBBHHHIiii", ref_id, pos, len(name), mapq, 4680, len(cigar), flag, len(seq), -1, -1, 0)
    body += name + cig + sq + b"\xff" * len(seq) + tags
    return struct.pack("<i", len(body)) + body


def test_spec_built_fixture(tmp_path):
    """IUPAC codes and '=' in SEQ count as N, odd lengths, CG:B,I long CIGAR, adjacent D D merged, a BGZF header with an
    extra subfield before BC, records spanning block boundaries, an empty block in the middle, unplaced reads last."""
    import struct
    hdr_text = b"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:cA\tLN:1000\n@SQ\tSN:cB\tLN:500\n"
    raw = b"BAM\x01" + struct.pack("<i", len(hdr_text)) + hdr_text + struct.pack("<i", 2)
    for nm, ln in ((b"cA\0", 1000), (b"cB\0", 500)):
        raw += struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln)
    r1 = _bam_record(0, 10, 60, 0, [(3, "S"), (4, "M"), (1, "D"), (2, "D"), (2, "="), (1, "X")], "ACGTRYAC=N")     # 10 bases
    real = [(2, "M"), (1, "I"), (4, "M")]                                                                      # 7 bases, stored in CG
    cgtag = b"CGBI" + struct.pack("<I", len(real)) + b"".join(struct.pack("<I", (n << 4) | "MIDNSHP=X".index(o)) for n, o in real)
    r2 = _bam_record(0, 40, 30, 16, [(7, "S"), (6, "N")], "ACGTACG", tags=b"NMCi\x01XZZhello\0" + cgtag)
    r3 = _bam_record(1, 5, 7, 1024, [(5, "M")], "TTTTT")
    r4 = _bam_record(-1, -1, 0, 4, [], "ACG")
    recs = r1 + r2 + r3 + r4
    stream = raw + recs
    cut1, cut2 = len(raw) + 17, len(raw) + len(r1) + 9              # block boundaries inside records
    path = tmp_path / "spec.bam"
    path.write_bytes(_bgzf_block(stream[:cut1], extra_first=struct.pack("<BBH", 88, 89, 3) + b"abc") + _bgzf_block(b"")
                     + _bgzf_block(stream[cut1:cut2]) + _bgzf_block(stream[cut2:]) + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    refs, got = read_bam(str(path))
    assert refs == [("cA", 1000), ("cB", 500)] and set(got) == {"cA", "cB"}
    a = got["cA"]
    assert list(a.pos) == [10, 40] and list(a.flag) == [0, 16] and list(a.mapq) == [60, 30]
    def cig(rd, i):
        return [(int(c) >> 4, "MIDNSHP=X"[int(c) & 15]) for c in rd.cigar[rd.cigar_off[i]:rd.cigar_off[i + 1]]]
    assert cig(a, 0) == [(3, "S"), (4, "M"), (3, "D"), (2, "="), (1, "X")]          # 1D2D merged
    assert cig(a, 1) == real                                                        # from the CG tag
    def bases(rd, i, n):
        k = int(rd.seq_off[i]) + np.arange(n)
        c = (rd.seq2[k >> 2] >> (2 * (k & 3))) & 3
        isn = ((rd.nmask[k >> 3] >> (k & 7)) & 1).astype(bool) if rd.nmask is not None else np.zeros(n, bool)
        return "".join("N" if m else "ACGT"[b] for b, m in zip(c, isn))
    assert bases(a, 0, 10) == "ACGTNNACNN" and bases(a, 1, 7) == "ACGTACG"
    assert (a.seq_off % 16 == 0).all()
    # padding after the odd-length read is clean (no N flag leaks from the pad nibble)
    assert bases(a, 1, 16)[7:] == "A" * 9
    b = got["cB"]
    assert list(b.pos) == [5] and list(b.flag) == [1024] and cig(b, 0) == [(5, "M")] and b.nmask is None


def test_index_fetch_and_contig_skip(tmp_path):
    """With a .bai: fetch(ref, beg, end) returns exactly the reads that overlap / start in the window, and
    contigs(only=...) seeks past unwanted references without inflating them."""
    from nanosnp_b200.bam import BamReader
    from nanosnp_b200.reads import max_reference_span
    cfg = SynthConfig(contig_len=400_000, coverage=8.0, len_median=3000, len_min=200, nbase_rate=0.002)
    _, rd1 = generate_host(cfg)
    cfg2 = SynthConfig(contig_len=150_000, coverage=6.0, len_median=2000, len_min=200, seed_reads=9, contig="ctg2")
    _, rd2 = generate_host(cfg2)
    path = str(tmp_path / "i.bam")
    write_bam(path, [("ctg1", 400_000), ("mid", 10), ("ctg2", 150_000)], {"ctg1": rd1, "ctg2": rd2}, index=True)
    with BamReader(path, threads=4) as r:
        assert r.has_index
        whole = {name: rd for _, name, rd in r.contigs()}
        total = r.inflated_bytes
    _same(rd1, whole["ctg1"]); _same(rd2, whole["ctg2"])
    ops = rd1.cigar & 15
    rl = np.where((ops == 0) | (ops == 2) | (ops == 3) | (ops == 7) | (ops == 8), rd1.cigar >> 4, 0).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(rl)])
    ends = rd1.pos + np.maximum(cs[rd1.cigar_off[1:]] - cs[rd1.cigar_off[:-1]], 1)
    for beg, end in ((0, 50_000), (123_456, 180_000), (390_000, 400_000), (200_000, 200_001)):
        with BamReader(path) as r:
            got = r.fetch(0, beg, end)
            part = r.inflated_bytes
        sel = np.nonzero((rd1.pos < end) & (ends > beg))[0]
        assert np.array_equal(got.pos, rd1.pos[sel]), (beg, end)
        assert np.array_equal(np.diff(got.cigar_off), np.diff(rd1.cigar_off)[sel])
        assert end - beg > 20_000 or part < total            # a short window only inflates its own byte range (+ read-ahead)
    with BamReader(path) as r:
        only = [(name, rd.n_reads) for _, name, rd in r.contigs({"ctg2"})]
        assert only == [("ctg2", rd2.n_reads)] and r.inflated_bytes < 0.6 * total
    with BamReader(path) as r:
        assert r.fetch(1, 0, 10).n_reads == 0


def test_spec_built_aux_tags_hp_and_qualities(tmp_path):
    """keep_aux: the HP tag is found behind every other aux type of the SAM spec (A c C s S i I f Z H B) and in every integer
    encoding htslib may choose for it; qualities come back at the reads' base offsets; the name hash is FNV-1a of QNAME."""
    import struct
    from nanosnp_b200.bam import BamReader, qname_hash
    hdr_text = b"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:cA\tLN:1000\n"
    raw = b"BAM\x01" + struct.pack("<i", len(hdr_text)) + hdr_text + struct.pack("<i", 1) + struct.pack("<i", 3) + b"cA\0" + struct.pack("<i", 1000)
    junk = (b"XAAq" + b"Xcc\xfe" + b"XCC\x07" + b"Xss" + struct.pack("<h", -300) + b"XSS" + struct.pack("<H", 60000) + b"Xii" + struct.pack("<i", -7)
            + b"XII" + struct.pack("<I", 4000000000) + b"Xff" + struct.pack("<f", 1.5) + b"XZZHP:i:2 not a tag\0" + b"XHH1AE3\0"
            + b"XBBs" + struct.pack("<I", 3) + struct.pack("<hhh", 1, 2, 3) + b"XDBC" + struct.pack("<I", 2) + b"HP")
    cases = [(b"HPC\x01", 1), (b"HPc\x02", 2), (b"HPS" + struct.pack("<H", 2), 2), (b"HPs" + struct.pack("<h", 1), 1),
             (b"HPi" + struct.pack("<i", 2), 2), (b"HPI" + struct.pack("<I", 1), 1), (b"", 0), (b"PSi" + struct.pack("<i", 77), 0)]
    recs, names, quals = b"", [], []
    for k, (tag, _) in enumerate(cases):
        name = b"read/%d\0" % k
        seq = "ACGTACGTA"[: 5 + (k % 5)]
        q = bytes((7 * k + j) % 60 for j in range(len(seq)))
        body = _bam_record(0, 10 + k, 60, 0, [(len(seq), "M")], seq, tags=(junk + tag) if k % 2 == 0 else (tag + junk), name=name)
        # _bam_record writes 0xFF qualities: patch the real ones in (they follow the packed sequence)
        off = 4 + 32 + len(name) + 4 + (len(seq) + 1) // 2
        body = body[:off] + q + body[off + len(seq):]
        recs += body; names.append(name[:-1].decode()); quals.append(q)
    path = tmp_path / "aux.bam"
    path.write_bytes(_bgzf_block(raw + recs) + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    with BamReader(str(path), keep_aux=True) as r:
        (_, _, rd), = list(r.contigs())
        ax = r.aux
    assert list(ax.hp) == [v for _, v in cases]
    assert list(ax.qhash) == [qname_hash(n) for n in names]
    for i, q in enumerate(quals):
        so = int(rd.seq_off[i])
        assert bytes(ax.qual[so:so + len(q)]) == q


def test_damaged_files_fail_cleanly(tmp_path):
    """Truncated / corrupted BGZF and BAM input must come back as errors (ValueError with the library's message), never as
    a crash or as silently shortened data: every prefix length of a small valid file, plus flipped bytes in the payload."""
    import struct
    from nanosnp_b200.bam import BamReader
    hdr_text = b"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:cA\tLN:1000\n"
    raw = b"BAM\x01" + struct.pack("<i", len(hdr_text)) + hdr_text + struct.pack("<i", 1) + struct.pack("<i", 3) + b"cA\0" + struct.pack("<i", 1000)
    recs = b"".join(_bam_record(0, 10 + 7 * k, 60, 0, [(4, "M"), (1, "I"), (5, "M")], "ACGTACGTAC", name=b"r%d\0" % k) for k in range(40))
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    cut = len(raw) + 500
    good = _bgzf_block((raw + recs)[:cut]) + _bgzf_block((raw + recs)[cut:]) + eof
    path = tmp_path / "ok.bam"
    path.write_bytes(good)
    with BamReader(str(path)) as r:
        (_, _, rd), = list(r.contigs())
    assert rd.n_reads == 40

    def outcome(data: bytes):
        p = tmp_path / "bad.bam"
        p.write_bytes(data)
        try:
            with BamReader(str(p)) as r:
                got = list(r.contigs())
            return sum(x[2].n_reads for x in got)
        except ValueError as e:
            assert str(e)                                   # the library's message, not an empty error
            return "error"
    results = {}
    for n in list(range(0, 120, 7)) + list(range(120, len(good) - len(eof), 37)) + [len(good) - len(eof) - 1]:
        results[n] = outcome(good[:n])
    assert all(v == "error" or (isinstance(v, int) and v <= 40) for v in results.values())
    assert results[0] == "error" and results[63] == "error"
    # a complete first block but a torn second one: the records of the first block may be returned only together with an error
    torn = outcome(good[: len(good) - len(eof) - 30])
    assert torn == "error"
    # corrupted deflate payload / wrong magic
    bad = bytearray(good); bad[40] ^= 0xFF; bad[41] ^= 0x55
    assert outcome(bytes(bad)) == "error"
    bad = bytearray(good); bad[0] = 0
    assert outcome(bytes(bad)) == "error"
    # a record whose block_size runs past the end of the stream
    lying = raw + struct.pack("<i", 10_000) + recs[4:]
    assert outcome(_bgzf_block(lying) + eof) == "error"
