"""VCF record text: the device formatter (csrc/vcf_dev.cu) against the host formatter (csrc/vcf.cu).

CPU: the host twin of the device code (same source, nsnp_vcf_format_records_at) gives the bytes of
nsnp_vcf_format_contig_records on random records, for batch-aligned ranges and -- through a batch-head table -- for arbitrary
ranges.  GPU: the kernels give the same bytes, streamed region by region with a carried partial batch, incl. the
rounding-tie fix-up."""
import ctypes as C

import numpy as np
import pytest

from nanosnp_b200 import _lib
from nanosnp_b200.predict_io import RECORD_DTYPE, REC_DROP, REC_TIE_GT, REC_TIE_ZY, AF_NAN, AF_ONE


def random_records(n, seed=0):
    rng = np.random.default_rng(seed)
    rec = np.zeros(n, RECORD_DTYPE)
    rec["gt"] = rng.choice(21, n, p=[0.08] * 10 + [0.2 / 11] * 11)
    rec["zy"] = rng.integers(0, 3, n)
    rec["ref"] = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
    rec["pos1"] = np.sort(rng.integers(1, 2_000_000_000, n)).astype(np.int32)
    rec["p_gt"] = rng.uniform(0.3, 1.0, n).astype(np.float32)
    rec["p_zy"] = rng.uniform(0.3, 1.0, n).astype(np.float32)
    from math import log
    k = -10.0 * (1.0 / log(10.0))
    for name, p in (("q100_gt", rec["p_gt"]), ("q100_zy", rec["p_zy"])):
        x = ((np.float32(1.0) - p) / p).astype(np.float64)
        with np.errstate(all="ignore"):
            t = np.maximum(k * np.log(np.maximum(x, 1e-300)) + 10.0, 0.0)
        rec[name] = np.floor(t * 100.0 + 0.5).astype(np.int64).clip(0, 2_000_000_000)
    rec["depth"] = rng.integers(-3, 200, n)
    rec["af_q"] = rng.integers(0, 1_000_001, n)
    rec["af_q"][rng.random(n) < 0.02] = AF_ONE
    rec["af_q"][rng.random(n) < 0.02] = AF_NAN
    fl = np.zeros(n, np.uint8)
    fl[rng.random(n) < 0.01] |= REC_DROP
    fl[rng.random(n) < 0.01] |= REC_TIE_GT
    fl[rng.random(n) < 0.01] |= REC_TIE_ZY
    rec["flags"] = fl
    return rec


def host_contig_text(lib, contig, rec, batch=1000):
    cap = len(rec) * 160 + 4096
    buf = C.create_string_buffer(cap)
    n = lib.nsnp_vcf_format_contig_records(contig.encode(), len(rec), rec.ctypes.data, batch, 4, C.addressof(buf), cap)
    assert n >= 0
    return buf.raw[:n]


def heads_table(rec, batch=1000):
    nb = (len(rec) + batch - 1) // batch
    t = np.full((nb, 10), 255, np.uint8)
    for b in range(nb):
        h = rec["gt"][b * batch:b * batch + 10]
        t[b, :len(h)] = h
    return t


def test_host_twin_matches_host_formatter():
    lib = _lib.load()
    for n, seed in ((25_317, 1), (1000, 2), (1007, 3), (9, 4), (1, 5)):
        rec = random_records(n, seed)
        want = host_contig_text(lib, "chr7", rec)
        cap = n * 160 + 4096
        buf = C.create_string_buffer(cap)
        w = lib.nsnp_vcf_format_records_at(b"chr7", rec.ctypes.data, n, 0, 1000, 0, C.addressof(buf), cap)
        assert buf.raw[:w] == want
        # arbitrary ranges through the batch-head table
        tab = heads_table(rec)
        cuts = sorted(set([0, n] + [int(x) for x in np.random.default_rng(seed).integers(0, n + 1, 5)]))
        got = b""
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = np.ascontiguousarray(rec[a:b])
            w = lib.nsnp_vcf_format_records_at(b"chr7", part.ctypes.data, b - a, a, 1000, tab.ctypes.data, C.addressof(buf), cap)
            got += buf.raw[:w]
        assert got == want
    assert want.count(b"\n") <= 1          # n = 1: at most one record


@pytest.mark.gpu
def test_gpu_text_streaming_table_and_ties():
    import torch
    from nanosnp_b200.vcf_text import GpuVcfText
    lib = _lib.load()
    dev = torch.device("cuda:0")
    n = 523_817
    rec = random_records(n, 11)
    rec["flags"] &= ~np.uint8(REC_TIE_GT | REC_TIE_ZY)
    # rounding ties: flagged records whose device digits are right, and flagged records whose device digits are WRONG (the
    # host fix-up must repair them, also when the repaired text is shorter / longer: x.x <-> x.xy)
    tie = np.random.default_rng(2).choice(n, 300, replace=False)
    rec["flags"][tie[:150]] |= REC_TIE_GT
    rec["flags"][tie[150:]] |= REC_TIE_ZY
    bad = tie[::3]
    rec["q100_gt"][bad] += 7
    rec["q100_zy"][bad] += 13
    want = host_contig_text(lib, "chr12", rec)
    rd = torch.from_numpy(rec.view(np.uint8).reshape(n, 32)).to(dev)
    # streaming: uneven region sizes, a region smaller than a batch, carried partial batches
    g = GpuVcfText(dev, "chr12", 1000)
    cuts = [0, 120_455, 120_999, 121_000, 121_650, 300_001, n]
    got = bytearray()
    for a, b in zip(cuts[:-1], cuts[1:]):
        ch = g.push(rd[a:b])
        if ch is not None:
            got += g.fetch(ch)
    ch = g.flush()
    if ch is not None:
        got += g.fetch(ch)
    assert len(got) == len(want)
    assert bytes(got) == want
    # arbitrary ranges with the batch-head table (multi-GPU form); the table is built by the device kernel too
    nb = (n + 999) // 1000
    heads = torch.full((nb, 10), 255, dtype=torch.uint8, device=dev)
    g2 = GpuVcfText(dev, "chr12", 1000)
    for a, b in zip(cuts[:-1], cuts[1:]):
        g2.batch_heads(rd[a:b], a, heads)
    assert np.array_equal(heads.cpu().numpy(), heads_table(rec))
    got2 = bytearray()
    for a, b in zip(cuts[:-1], cuts[1:]):
        got2 += g2.fetch(g2.format_at(rd[a:b], a, heads))
    assert bytes(got2) == want
    # empty input and a contig shorter than ten sites
    g3 = GpuVcfText(dev, "c", 1000)
    assert g3.push(rd[:0]) is None and g3.flush() is None
    g4 = GpuVcfText(dev, "c", 1000)
    assert g4.push(rd[:7]) is None
    small = bytes(g4.fetch(g4.flush()))
    assert small == host_contig_text(lib, "c", np.ascontiguousarray(rec[:7]))


@pytest.mark.gpu
def test_gpu_sharded_writer_streaming_equals_batch_and_host(tmp_path):
    """ShardedVcfWriter: begin / add_region / finish (deferred batch heads, text copied out region by region) writes the same
    file as write() and as the host contig formatter, incl. a contig whose last batch has fewer than ten sites (dropped fix-up
    records) and rounding ties."""
    import torch
    from nanosnp_b200.caller import ShardedVcfWriter
    from nanosnp_b200.shard import plan_regions
    lib = _lib.load()
    dev = torch.device("cuda:0")
    contigs = [("cA", 400_000), ("cB", 90_000), ("cC", 150_000)]
    regions = plan_regions(contigs, 50_000)
    recs, want = {}, b"##h\n"
    for ci, (name, L) in enumerate(contigs):
        n = {0: 61_234, 1: 9_003, 2: 20_000}[ci]              # cB: the last batch holds 3 sites
        rec = random_records(n, 100 + ci)
        rec["zy"][-3:] = 1; rec["gt"][-3:] = 0; rec["ref"][-3:] = ord("A"); rec["flags"][-3:] = 0     # fix-up records in the short batch
        rec["pos1"] = np.sort(np.random.default_rng(ci).choice(L, n, replace=False)).astype(np.int32) + 1
        rec["flags"] &= ~np.uint8(REC_TIE_GT | REC_TIE_ZY)
        tie = np.random.default_rng(9 + ci).choice(n, 60, replace=False)
        rec["flags"][tie] |= REC_TIE_GT
        rec["q100_gt"][tie[::2]] += 5
        want += host_contig_text(lib, name, rec)
        for i, rg in enumerate(regions):
            if rg.contig_index == ci:
                sel = (rec["pos1"] - 1 >= rg.emit_start) & (rec["pos1"] - 1 < rg.emit_end)
                recs[i] = torch.from_numpy(np.ascontiguousarray(rec[sel]).view(np.uint8).reshape(-1, 32)).to(dev)
    w = ShardedVcfWriter(contigs, regions, 1000, dev)
    p1 = str(tmp_path / "batch.vcf")
    w.write(p1, b"##h\n", recs)
    assert open(p1, "rb").read() == want
    for rep in range(2):                                        # twice: buffers are reused
        p2 = str(tmp_path / f"stream{rep}.vcf")
        w.begin()
        for i in sorted(recs):
            w.add_region(i, recs[i])
        info = w.finish(p2, b"##h\n")
        assert open(p2, "rb").read() == want and info["sites"] == sum(int(r.shape[0]) for r in recs.values())
