"""Pins the oracle (oracle/*.c, oracle/s2_restate.py) to the reference:
   - committed golden files produced by the reference's own binaries / Python (tests/golden/make_golden.py)
   - live runs of oracle/_ref when those binaries are present (build container and GPU box snapshot)."""
import filecmp
import os

import numpy as np
import pytest

from conftest import oracle_s1


def _restate_text(orc, golden, tmp_path, stem, ref_name, **kw):
    ref = np.frombuffer((golden / ref_name).read_bytes(), np.uint8)
    out_t, out_p = str(tmp_path / "o.tensor"), str(tmp_path / "o.pd")
    res = orc.s1_restate(str(golden / f"{stem}.mpileup"), "ctg1", ref, tensor_path=out_t, pd_path=out_p, **kw)
    return res, out_t, out_p


def test_appendix_c1(orc, golden, tmp_path):
    res, t, p = _restate_text(orc, golden, tmp_path, "c1", "c1.ref")
    assert list(res.positions) == [40, 45]
    assert filecmp.cmp(t, golden / "c1.tensor", shallow=False)
    assert filecmp.cmp(p, golden / "c1.pd", shallow=False)
    # SURVEY appendix C-1 known answers
    assert list(res.windows[0, 16]) == [-5, 2, 0, 0, 0, 0, 0, 0, 0, -5, 2, 0, 0, 0, 0, 0, 0, 0]
    assert list(res.windows[1, 16]) == [0, 0, 0, -8, 0, 0, 2, 2, 0, 0, 0, 0, -5, 2, 2, 0, 0, 0]
    assert list(res.windows[1, 17]) == [0, -3, 0, 0, 0, 0, 0, 0, 2, 0, -5, 0, 0, 0, 0, 0, 0, 0]
    assert list(res.depth) == [10, 13]


@pytest.mark.parametrize("tag,kw,expect", [("c2_af012", {}, [60, 70]), ("c2_af09", {"snp_min_af": 0.9, "indel_min_af": 0.9}, [60])])
def test_appendix_c2(orc, golden, tmp_path, tag, kw, expect):
    res, t, _ = _restate_text(orc, golden, tmp_path, "c2", "c2.ref", **kw)
    assert list(res.positions) == expect
    assert filecmp.cmp(t, golden / f"{tag}.tensor", shallow=False)


def test_small_case_matches_reference_binaries(orc, small_case, tmp_path):
    res = oracle_s1(orc, small_case["reads"], small_case["ref"], tmp_path)
    assert np.array_equal(res.positions, small_case["site_pos"])
    assert np.array_equal(res.windows, small_case["windows"])
    assert len(res.positions) > 3000
    # window rows are rows of the count array
    c = res.counts
    for k in (0, 17, len(res.positions) - 1):
        p = res.positions[k] - 1
        assert np.array_equal(res.windows[k], c[p - 16:p + 17])


@pytest.mark.parametrize("seed,kw", [(1, {}), (2, {"use_eqx": True, "nbase_rate": 0.01}), (3, {"coverage": 45.0, "gap_period": 5000, "gap_len": 120})])
def test_live_against_ref_binaries(orc, tmp_path, seed, kw):
    if not orc.have_ref_binaries():
        pytest.skip("oracle/_ref binaries not built")
    from nanosnp_b200.synth import SynthConfig, generate_host
    base = dict(contig_len=30_000, coverage=25.0, seed_ref=seed, seed_var=seed + 50, seed_reads=seed + 90, len_median=2500,
                len_min=200, ref_n_period=7000, ref_n_len=30, ref_lower_period=2900, ref_lower_len=150)
    base.update(kw)
    cfg = SynthConfig(**base)
    ref, reads = generate_host(cfg)
    os.makedirs(tmp_path / "pile")
    mp = str(tmp_path / "pile" / "ctg1.mpileup")
    orc.mpileup_text(reads, "ctg1", mp)
    orc.write_fasta(str(tmp_path / "ref.fa"), {"ctg1": ref})
    t_ref, p_ref = orc.s1_reference(mp, str(tmp_path / "ref.fa"), "ctg1", str(tmp_path))
    res = orc.s1_restate(mp, "ctg1", ref, tensor_path=str(tmp_path / "o.tensor"), pd_path=str(tmp_path / "o.pd"))
    assert len(res.positions) > 500
    assert filecmp.cmp(t_ref, tmp_path / "o.tensor", shallow=False)
    assert filecmp.cmp(p_ref, tmp_path / "o.pd", shallow=False)
    x, _, pos, refb = orc.parse_pd(p_ref)
    assert np.array_equal(x, res.windows) and np.array_equal(pos, res.positions)


def test_model_oracle_matches_reference_python(golden, golden_weights, small_case):
    from oracle.s2_restate import PileupModelOracle
    import torch
    torch.set_num_threads(1)
    m = PileupModelOracle(*golden_weights)
    gt, zy = m.predict(small_case["windows"])
    z = np.load(golden / "s2_small.npz")
    assert np.abs(gt.numpy() - z["gt"]).max() < 2e-6 and np.abs(zy.numpy() - z["zy"]).max() < 2e-6
    assert np.array_equal(gt.numpy().argmax(1), z["gt"].argmax(1))
    # SURVEY appendix C-3
    x = np.zeros((2, 33, 18), np.float32); x[:, :, 0] = -10; x[:, :, 9] = -10; x[1, 16, 1] = 5; x[1, 16, 10] = 5
    g, y = m.predict(x)
    assert g.argmax(1).tolist() == [0, 1] and y.argmax(1).tolist() == [0, 2]
    assert abs(float(g[0, 0]) - 0.8981) < 1e-3 and abs(float(y[1, 2]) - 0.4939) < 1e-3


def test_vcf_restatement_matches_reference_python(golden, small_case):
    from oracle.s2_restate import vcf_header, vcf_records
    z = np.load(golden / "s2_small.npz")
    n = len(small_case["site_pos"])
    text = vcf_header(open(golden / "s2_small.fai").read().splitlines())
    for b in range(0, n, 1000):
        sl = slice(b, min(n, b + 1000))
        text += vcf_records(["ctg1"] * (sl.stop - sl.start), small_case["site_pos"][sl].astype(np.int64),
                            small_case["site_refbase"][sl].astype(np.int64), small_case["windows"][sl].astype(np.float32),
                            z["gt"][sl], z["zy"][sl])
    assert text == (golden / "s2_small.vcf").read_text()
    # batch size 7: fewer than 10 sites per batch -> IndexError quirk drops / rewrites records
    text = vcf_header(open(golden / "s2_small.fai").read().splitlines())
    for b in range(0, 61, 7):
        sl = slice(b, min(61, b + 7))
        text += vcf_records(["ctg1"] * (sl.stop - sl.start), small_case["site_pos"][sl].astype(np.int64),
                            small_case["site_refbase"][sl].astype(np.int64), small_case["windows"][sl].astype(np.float32),
                            z["gt"][sl], z["zy"][sl])
    assert text == (golden / "s2_tiny_b7.vcf").read_text()
