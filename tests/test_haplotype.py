"""BASELINE configs[4] (HaplotypeModel s5).  The shipped haplotype checkpoints are missing, so parity is
against the REAL reference Python run with seeded random-init weights (tests/golden/make_golden_hap.py).

CPU: the oracle restatement reproduces the reference's features, probabilities and CSV.  GPU: the CUDA feature kernel is
bit-exact against the oracle, the fp32 CUDA model is within 2e-5 of a float64 evaluation of the oracle network, argmax equal,
CSV identical."""
import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def hap():
    return np.load(GOLDEN / "hap_small.npz")


def test_oracle_matches_reference_python(hap):
    import torch
    from oracle import hap_restate as H
    torch.set_num_threads(1)
    for name, L in (("pileup", 33), ("haplotype", 11)):
        got = np.stack([H.frequency_features(hap[f"{name}_seq"][i], hap[f"{name}_bq"][i], hap[f"{name}_mq"][i], hap[f"{name}_hp"][i])
                        for i in range(len(hap["pos"]))])
        assert np.array_equal(got, hap[f"{name}_feat"][:, :104])             # float64, bit for bit
    refrow = np.stack([H.reference_codes(hap["ref"], np.arange(p - 16, p + 17)) for p in hap["pos"]])
    assert np.array_equal(refrow, hap["pileup_feat"][:, 104].astype(np.int64))
    assert np.array_equal(np.stack([H.reference_codes(hap["ref"], q) for q in hap["hap_pos"]]), hap["haplotype_feat"][:, 104].astype(np.int64))
    m = H.HaplotypeModelOracle(seed=int(hap["seed"]))
    assert sum(p.numel() for p in m.parameters()) == int(hap["n_params"]) == 8192013
    gt, zy = m.predict(hap["pileup_feat"], hap["haplotype_feat"])
    assert np.abs(gt.numpy() - hap["gt"]).max() < 1e-6 and np.abs(zy.numpy() - hap["zy"]).max() < 1e-6
    rows = H.predict_rows(["ctgH:%d" % p for p in hap["pos"]], hap["gt"])
    assert rows == (GOLDEN / "hap_small.csv").read_text()


def test_state_dict_contract():
    from nanosnp_b200.haplotype import LSTMNetwork, required_keys
    from oracle.hap_restate import HaplotypeModelOracle
    m = HaplotypeModelOracle(seed=1)
    assert sorted(required_keys()) == sorted(m.state_dict().keys())          # the reference checkpoint's key set
    net = LSTMNetwork()
    with pytest.raises(RuntimeError):
        net.load_state_dict({"forward_layer.dense.weight": m.state_dict()["forward_layer.dense.weight"]})
    net.load_state_dict(m.state_dict())
    with pytest.raises(Exception):
        net.predict(None, None)                                              # no device: there is no CPU fallback


@pytest.mark.gpu
def test_gpu_features_bit_exact_and_model_within_tolerance(hap, tmp_path):
    import torch
    from nanosnp_b200 import haplotype as G
    from oracle import hap_restate as H
    n = len(hap["pos"])
    # ---- H4: features, bit-exact after the reference's own float32 cast ----
    pref = np.stack([G.reference_codes(hap["ref"], np.arange(p - 16, p + 17)) for p in hap["pos"]])
    href = np.stack([G.reference_codes(hap["ref"], q) for q in hap["hap_pos"]])
    fp = G.frequency_features(hap["pileup_seq"], hap["pileup_bq"], hap["pileup_mq"], hap["pileup_hp"], pref)
    fh = G.frequency_features(hap["haplotype_seq"], hap["haplotype_bq"], hap["haplotype_mq"], hap["haplotype_hp"], href)
    assert np.array_equal(fp.cpu().numpy(), hap["pileup_feat"].astype(np.float32))
    assert np.array_equal(fh.cpu().numpy(), hap["haplotype_feat"].astype(np.float32))
    # ---- H5: model (seeded random-init weights: the shipped checkpoints are missing) ----
    m = H.HaplotypeModelOracle(seed=int(hap["seed"]))
    net = G.LSTMNetwork().to("cuda")
    net.load_state_dict(m.state_dict())
    gt, zy = net.predict(fp, fh)
    g64, z64 = m.predict(hap["pileup_feat"], hap["haplotype_feat"], dtype=torch.float64)
    err = max(float((gt.cpu() - g64).abs().max()), float((zy.cpu() - z64).abs().max()))
    assert err < 2e-5, err
    assert np.abs(gt.cpu().numpy() - hap["gt"]).max() < 2e-5 and np.abs(zy.cpu().numpy() - hap["zy"]).max() < 2e-5     # the reference's own output
    assert torch.equal(gt.argmax(1).cpu(), g64.argmax(1))
    # ragged batch sizes around the 64-site tile and more than one 2048-site chunk
    big_p = fp.repeat(50, 1, 1)[:2100]; big_h = fh.repeat(50, 1, 1)[:2100]
    g2, z2 = net.predict(big_p, big_h)
    assert torch.equal(g2[:n], gt) and torch.equal(g2[n:2 * n], gt) and torch.equal(g2[2064:2100], gt[2064 % n:2064 % n + 36])
    for k in (1, 63, 65):
        gk, zk = net.predict(fp[:k], fh[:k])
        assert torch.equal(gk, gt[:k]) and torch.equal(zk, zy[:k])
    # ---- H6: the predict loop over an .npz of write_to_bins.py's arrays -> CSV identical to the reference's ----
    d = tmp_path / "bins"; d.mkdir()
    np.savez(d / "ctgH_1_5000.npz", pileup_sequences=hap["pileup_seq"], pileup_hap=hap["pileup_hp"], pileup_baseq=hap["pileup_bq"],
             pileup_mapq=hap["pileup_mq"], haplotype_sequences=hap["haplotype_seq"], haplotype_hap=hap["haplotype_hp"],
             haplotype_baseq=hap["haplotype_bq"], haplotype_mapq=hap["haplotype_mq"],
             candidate_positions=np.array(["ctgH:%d" % p for p in hap["pos"]]),
             haplotype_positions=np.array([["ctgH:%d" % q for q in row] for row in hap["hap_pos"]]))
    fa = tmp_path / "ref.fa"
    fa.write_bytes(b">ctgH some description\n" + bytes(hap["ref"]) + b"\n")
    out = tmp_path / "haplotype.csv"
    rows = G.predict(net, str(d), str(fa), 20, 33, 11, str(out), "cuda")
    assert rows == n and out.read_text() == (GOLDEN / "hap_small.csv").read_text()
