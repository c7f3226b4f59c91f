"""CPU-side checks of the product library: the C-ABI shared object loads and exports every declared symbol, the
host-only entry points (VCF formatter, synthetic generator) behave, and the GPU entry points refuse loudly
instead of falling back when no device is present."""
import ctypes as C
import re

import numpy as np
import pytest

from nanosnp_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = (_lib._HERE.parent / "include" / "nanosnp_b200.h").read_text()
    declared = set(re.findall(r"\b(nsnp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.nsnp_abi_version() == 2
    p = _lib.default_params()
    assert (p.snp_min_af, p.indel_min_af, p.min_coverage, p.min_mapq, p.excl_flags) == (0.12, 0.12, 6, 20, 2316)
    assert lib.nsnp_model_blob_bytes() > 700_000
    assert lib.nsnp_pileup_workspace_bytes(1000, 100000, 1 << 20) > 0


def test_no_cpu_fallback():
    import torch
    lib = _lib.load()
    if lib.nsnp_device_count() > 0 and torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nanosnp_b200.pipeline import PileupEngine
    with pytest.raises(_lib.NsnpError) as ei:
        PileupEngine("cuda:0")
    assert ei.value.code == _lib.E_NO_DEVICE
    with pytest.raises(_lib.NsnpError):
        PileupEngine("cpu")
    # the raw C entry points refuse as well
    p = _lib.default_params()
    r = _lib.Reads()
    buf = (C.c_char * 4096)()
    a = C.addressof(buf)
    rc = lib.nsnp_pileup_counts(C.byref(r), a, 100, 0, 100, C.byref(p), a, a, a, 4096, a, None)
    assert rc == _lib.E_NO_DEVICE, lib.nsnp_last_error()
    rc = lib.nsnp_pileup_model_forward(a, a, None, 4, None, a, a, a, 1 << 30, _lib.PREC_FP32, None)
    assert rc == _lib.E_NO_DEVICE


def _native_vcf(contig, pos, refb, gt, zy, x, batch):
    lib = _lib.load()
    out = []
    n = len(pos)
    cov = np.ascontiguousarray(x[:, 16, [0, 1, 2, 3, 9, 10, 11, 12]].astype(np.float32))
    for b in range(0, n, batch):
        e = min(n, b + batch)
        cap = (e - b) * 160 + 64
        buf = C.create_string_buffer(cap)
        p = np.ascontiguousarray(pos[b:e], np.int32); r = np.ascontiguousarray(refb[b:e], np.uint8)
        g = np.ascontiguousarray(gt[b:e], np.float32); z = np.ascontiguousarray(zy[b:e], np.float32)
        c = np.ascontiguousarray(cov[b:e])
        m = lib.nsnp_vcf_format_batch(contig.encode(), e - b, p.ctypes.data, r.ctypes.data, g.ctypes.data, z.ctypes.data,
                                      c.ctypes.data, C.addressof(buf), cap)
        assert m >= 0
        out.append(buf.raw[:m].decode())
    return "".join(out)


def test_native_vcf_formatter_matches_reference_python(golden, small_case):
    from oracle.s2_restate import vcf_header
    z = np.load(golden / "s2_small.npz")
    hdr = vcf_header(open(golden / "s2_small.fai").read().splitlines())
    body = _native_vcf("ctg1", small_case["site_pos"], small_case["site_refbase"], z["gt"], z["zy"], small_case["windows"], 1000)
    assert hdr + body == (golden / "s2_small.vcf").read_text()
    body = _native_vcf("ctg1", small_case["site_pos"][:61], small_case["site_refbase"][:61], z["gt"][:61], z["zy"][:61],
                       small_case["windows"][:61], 7)
    assert hdr + body == (golden / "s2_tiny_b7.vcf").read_text()


def test_fast_threaded_formatter_is_byte_identical_to_the_libc_one(golden, small_case):
    from nanosnp_b200.predict import format_records
    z = np.load(golden / "s2_small.npz")
    x = small_case["windows"]; cov = x[:, 16, [0, 1, 2, 3, 9, 10, 11, 12]].astype(np.float32)
    for batch, threads in ((1000, 1), (1000, 5), (7, 3), (64, 16)):
        slow = _native_vcf("ctg1", small_case["site_pos"], small_case["site_refbase"], z["gt"], z["zy"], x, batch)
        fast = format_records("ctg1", small_case["site_pos"], small_case["site_refbase"], z["gt"], z["zy"], cov, batch, threads).decode()
        assert fast == slow, (batch, threads)
    # adversarial numerics: probabilities on a dense grid (QUAL rounding ties), AF ties (n/128), p == 1, zero depth
    rng = np.random.default_rng(11)
    n = 20000
    gt = np.full((n, 21), 1e-9, np.float32); zy = np.full((n, 3), 1e-9, np.float32)
    pmax = np.linspace(0.05, 0.99999, n).astype(np.float32)
    k = rng.integers(0, 10, n); gt[np.arange(n), k] = pmax; gt[::501, :] = 0; gt[::501, 3] = 1.0
    kz = rng.integers(0, 3, n); zy[np.arange(n), kz] = rng.uniform(0.34, 1.0, n).astype(np.float32)
    xx = np.zeros((n, 33, 18), np.float32)
    refb = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)]
    for i in range(n):
        r = "ACGT".index(chr(refb[i])); d = int(rng.choice([0, 6, 64, 128, 100]))
        alt = (r + 1) % 4; s = int(rng.integers(0, d + 1)) if d else 0
        xx[i, 16, alt] = s; xx[i, 16, r] = -d
    cov = np.ascontiguousarray(xx[:, 16, [0, 1, 2, 3, 9, 10, 11, 12]])
    pos = np.arange(1, n + 1).astype(np.int32)
    slow = _native_vcf("chr7", pos, refb, gt, zy, xx, 1000)
    fast = format_records("chr7", pos, refb, gt, zy, cov, 1000, 4).decode()
    assert fast == slow
    assert "nan" in fast and ":1.000000" in fast


def test_compact_record_formatter_matches_the_array_formatter(golden, small_case):
    """Host half of the GPU record path: records computed by the NumPy statement of site_record_kernel must give the same
    bytes as the array formatter (golden set, tiny batches, adversarial ties / nan / p == 1)."""
    from nanosnp_b200.predict_io import (ContigVcfAssembler, format_compact_records_into, format_records, records_reference, vcf_buffer_bytes)
    z = np.load(golden / "s2_small.npz")
    x = small_case["windows"]; cov = np.ascontiguousarray(x[:, 16, [0, 1, 2, 3, 9, 10, 11, 12]].astype(np.float32))
    n = 2500
    rec = records_reference(small_case["site_pos"][:n], small_case["site_refbase"][:n], z["gt"][:n], z["zy"][:n], cov[:n])
    buf = np.empty(vcf_buffer_bytes(n, "ctg1"), np.uint8)
    for batch, th in ((1000, 3), (7, 2)):
        w = format_compact_records_into(buf, "ctg1", rec, batch, th)
        assert buf[:w].tobytes() == format_records("ctg1", small_case["site_pos"][:n], small_case["site_refbase"][:n], z["gt"][:n], z["zy"][:n], cov[:n], batch, th)
    rng = np.random.default_rng(12)
    m = 3000
    gt = np.full((m, 21), 1e-9, np.float32); zy = np.full((m, 3), 1e-9, np.float32)
    gt[np.arange(m), rng.integers(0, 12, m)] = np.linspace(0.05, 0.99999, m).astype(np.float32); gt[::301, :] = 0; gt[::301, 2] = 1.0
    zy[np.arange(m), rng.integers(0, 3, m)] = rng.uniform(0.34, 1.0, m).astype(np.float32)
    refb = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, m)]
    c8 = np.zeros((m, 8), np.float32)
    for i in range(m):
        r = "ACGT".index(chr(refb[i])); d = int(rng.choice([0, 6, 64, 128, 100])); a = (r + 1) % 4
        c8[i, a] = int(rng.integers(0, d + 1)) if d else 0; c8[i, r] = -d
    pos = np.arange(1, m + 1).astype(np.int32)
    rec = records_reference(pos, refb, gt, zy, c8)
    buf = np.empty(vcf_buffer_bytes(m, "chrQ"), np.uint8)
    w = format_compact_records_into(buf, "chrQ", rec, 1000, 4)
    assert buf[:w].tobytes() == format_records("chrQ", pos, refb, gt, zy, c8, 1000, 4)
    # streaming assembler in records mode
    import io
    sink = io.BytesIO(); asm = ContigVcfAssembler("chrQ", 1000, 2, sink)
    for a, b in ((0, 999), (999, 1001), (1001, 3000)):
        asm.add_records(rec[a:b])
    asm.close()
    assert sink.getvalue() == buf[:w].tobytes()


def test_contig_assembler_keeps_batch_composition(golden, small_case):
    """Region-by-region streaming must give the same text as formatting the whole contig at once."""
    import io
    from nanosnp_b200.predict import ContigVcfAssembler, format_records
    z = np.load(golden / "s2_small.npz")
    x = small_case["windows"]; cov = np.ascontiguousarray(x[:, 16, [0, 1, 2, 3, 9, 10, 11, 12]].astype(np.float32))
    pos0 = (small_case["site_pos"] - 1).astype(np.int32)
    whole = format_records("ctg1", small_case["site_pos"], small_case["site_refbase"], z["gt"], z["zy"], cov, 1000, 2)
    rng = np.random.default_rng(3)
    for trial in range(4):
        cuts = np.sort(rng.choice(np.arange(1, len(pos0)), size=6, replace=False)).tolist()
        if trial == 0:
            cuts = [1, 2, 999, 1000, 1001, 3000]
        sink = io.BytesIO()
        asm = ContigVcfAssembler("ctg1", 1000, 2, sink)
        for a, b in zip([0] + cuts, cuts + [len(pos0)]):
            asm.add(pos0[a:b], small_case["site_refbase"][a:b], z["gt"][a:b], z["zy"][a:b], cov[a:b])
        asm.close()
        assert sink.getvalue() == whole
    assert whole.decode() == "".join(l + "\n" for l in (golden / "s2_small.vcf").read_text().splitlines() if not l.startswith("#"))


def test_native_vcf_formatter_quirks_vs_oracle():
    """Randomised probabilities (incl. p == 1.0, tiny batches, zero depth) against the Python restatement."""
    from oracle.s2_restate import vcf_records
    rng = np.random.default_rng(5)
    for batch in (1, 3, 9, 10, 50):
        n = 400
        logits = rng.normal(size=(n, 21)).astype(np.float32) * 4
        logits[:, 10:] -= 3
        gt = np.exp(logits - logits.max(1, keepdims=True)); gt = (gt / gt.sum(1, keepdims=True)).astype(np.float32)
        lz = rng.normal(size=(n, 3)).astype(np.float32) * 3
        zy = np.exp(lz - lz.max(1, keepdims=True)); zy = (zy / zy.sum(1, keepdims=True)).astype(np.float32)
        gt[::37] = 0; gt[::37, 4] = 1.0          # p == 1.0 -> calculate_score raises -> record dropped
        x = rng.integers(0, 30, size=(n, 33, 18)).astype(np.float32)
        refb = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)]
        for i in range(n):
            k = "ACGT".index(chr(refb[i]))
            x[i, 16, k] = -x[i, 16, :4].sum(); x[i, 16, 9 + k] = -x[i, 16, 9:13].sum()
        x[5, 16, :] = 0                           # depth 0 -> nan AF
        pos = np.arange(100, 100 + n).astype(np.int32)
        exp = ""
        for b in range(0, n, batch):
            e = min(n, b + batch)
            exp += vcf_records(["chrZ"] * (e - b), pos[b:e].astype(np.int64), refb[b:e].astype(np.int64), x[b:e], gt[b:e], zy[b:e])
        got = _native_vcf("chrZ", pos, refb, gt, zy, x, batch)
        assert got == exp, batch


def test_synth_generator_properties():
    from nanosnp_b200.synth import SynthConfig, generate_host
    from nanosnp_b200.reads import reference_span
    cfg = SynthConfig(contig_len=50_000, coverage=12, nbase_rate=0.01, gap_period=20_000, gap_len=400, len_median=2000, len_min=150)
    ref, rd = generate_host(cfg)
    ref2, rd2 = generate_host(cfg)
    assert np.array_equal(ref, ref2) and np.array_equal(rd.cigar, rd2.cigar) and np.array_equal(rd.seq2, rd2.seq2)
    assert (np.diff(rd.pos) >= 0).all() and rd.n_reads > 100
    assert (rd.seq_off % 16 == 0).all()
    npass = 0
    for r in range(rd.n_reads):
        cg = rd.cigar[rd.cigar_off[r]:rd.cigar_off[r + 1]]
        ops = "".join("MIDNSHP=X"[o] for o in (cg & 15))
        # appendix B.4 subset: [S] M {(I|D) M}* [S]
        assert re.fullmatch(r"S?M((I|D)M)*S?", ops), ops
        assert rd.pos[r] + reference_span(cg) <= cfg.contig_len
        if not (rd.flag[r] & 2316) and rd.mapq[r] >= 20:
            npass += 1
            # no passing read overlaps a coverage gap
            s, e = int(rd.pos[r]), int(rd.pos[r]) + reference_span(cg)
            assert (s % 20_000) < 19_600 and ((e - 1) % 20_000) < 19_600 and (e - s) < 20_000
    assert npass > 0.9 * rd.n_reads * 0.9
    assert set(np.unique(ref)) <= set(b"ACGTacgtNnR")


def test_compact_record_formatter_edge_values():
    """Largest positions / depths, a long contig name, AF = 1 and nan, every digit count of QUAL: the hand-rolled decimal
    writer must print what Python prints."""
    from nanosnp_b200.predict_io import AF_NAN, AF_ONE, RECORD_DTYPE, format_compact_records_into, vcf_buffer_bytes
    name = "chrUn_" + "x" * 150
    pos = np.array([1, 9, 10, 99, 100, 12345, 999999, 1000000, 99999999, 100000000, 2147483647], np.int32)
    n = len(pos)
    rec = np.zeros(n, RECORD_DTYPE)
    rec["gt"] = 1; rec["zy"] = 2; rec["ref"] = ord("A"); rec["pos1"] = pos            # AC, 0/1, ref A -> ALT C
    rec["q100_gt"] = [0, 5, 10, 99, 100, 101, 1230, 9999, 10000, 123456, 30100]; rec["q100_zy"] = 500000
    rec["depth"] = [0, 6, 9, 10, 99, 100, 144, 1000, 65535, 16383, 2000000000]
    rec["af_q"] = [0, 1, 10, 999999, AF_ONE, AF_NAN, 500000, 123456, 100000, 99, 1000000]
    rec["p_gt"] = 0.5; rec["p_zy"] = 0.5
    buf = np.empty(vcf_buffer_bytes(n, name), np.uint8)
    w = format_compact_records_into(buf, name, rec, 1000, 1)
    lines = buf[:w].tobytes().decode().splitlines()
    assert len(lines) == n
    for i, l in enumerate(lines):
        q = int(rec["q100_gt"][i])
        qs = str(float(round(q / 100.0, 2)))
        af = "1.000000" if rec["af_q"][i] == AF_ONE else "nan" if rec["af_q"][i] == AF_NAN else "%f" % (int(rec["af_q"][i]) / 1e6)
        exp = "%s\t%d\t.\tA\tC\t%s\tPASS\t.\tGT:GQ:DP:AF\t0/1:%d:%d:%s" % (name, pos[i], qs, int(float(qs)), int(rec["depth"][i]), af)
        assert l == exp, (i, l, exp)


def test_checkpoint_fixture_loader_and_blob_size(golden):
    """The product-side .npz checkpoint loader returns the two state dicts of utils.py:67-77; packing them fills the blob
    the library sizes (host-only calls: no GPU needed)."""
    import ctypes as C
    from nanosnp_b200 import _lib
    from nanosnp_b200.utils import load_weights_npz
    enc, fwd = load_weights_npz(golden / "ont_pileup_weights.npz")
    assert set(enc) >= {"lstm.weight_ih_l0", "lstm.weight_hh_l1_reverse", "output_proj.weight"} and set(fwd) >= {"dense.weight", "genotype_layer.bias"}
    assert enc["lstm.weight_ih_l0"].shape == (256, 18) and enc["lstm.weight_ih_l1"].shape == (256, 128) and fwd["genotype_layer.weight"].shape == (21, 256)
    lib = _lib.load()
    nbytes = lib.nsnp_model_blob_bytes()
    assert nbytes > 1_000_000
    keep = []

    def ptr(a):
        a = np.ascontiguousarray(a, dtype=np.float32); keep.append(a); return a.ctypes.data
    w = _lib.ModelWeights()
    for layer in range(2):
        for d, sfx in enumerate(("", "_reverse")):
            i = layer * 2 + d
            w.w_ih[i] = ptr(enc[f"lstm.weight_ih_l{layer}{sfx}"]); w.w_hh[i] = ptr(enc[f"lstm.weight_hh_l{layer}{sfx}"])
            w.b_ih[i] = ptr(enc[f"lstm.bias_ih_l{layer}{sfx}"]); w.b_hh[i] = ptr(enc[f"lstm.bias_hh_l{layer}{sfx}"])
    w.proj_w, w.proj_b = ptr(enc["output_proj.weight"]), ptr(enc["output_proj.bias"])
    w.dense_w, w.dense_b = ptr(fwd["dense.weight"]), ptr(fwd["dense.bias"])
    w.gt_w, w.gt_b = ptr(fwd["genotype_layer.weight"]), ptr(fwd["genotype_layer.bias"])
    w.zy_w, w.zy_b = ptr(fwd["zygosity_layer.weight"]), ptr(fwd["zygosity_layer.bias"])
    blob = np.zeros(nbytes, np.uint8)
    assert lib.nsnp_model_pack_weights(C.byref(w), blob.ctypes.data, nbytes) == 0
    assert lib.nsnp_model_pack_weights(C.byref(w), blob.ctypes.data, nbytes - 1) < 0          # too small: refused, not truncated
    tail = blob[-2048:].view(np.float32)                                                     # layer-1 bias (scaled), last block of the blob
    b = (enc["lstm.bias_ih_l1_reverse"] + enc["lstm.bias_hh_l1_reverse"]).astype(np.float32)
    assert np.isclose(np.abs(tail[256:]).max(), max(np.abs(b[128:192]).max() * 2 * np.log2(np.e), np.abs(np.delete(b, np.s_[128:192])).max() * np.log2(np.e)), rtol=1e-5)
