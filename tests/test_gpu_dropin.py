"""The drop-in entry points on a GPU: `python -m nanosnp_b200.predict` with the reference's CLI must reproduce the
reference Python's VCF from (a) the reference's own .pd text hand-off and (b) packed reads (s1 on the GPU too)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare_vcf(got_text, ref_text):
    got = [l for l in got_text.splitlines() if not l.startswith("#")]
    ref = [l for l in ref_text.splitlines() if not l.startswith("#")]
    assert [l for l in got_text.splitlines() if l.startswith("#")] == [l for l in ref_text.splitlines() if l.startswith("#")]
    assert len(got) == len(ref)
    nq = 0
    for a, b in zip(got, ref):
        fa, fb = a.split("\t"), b.split("\t")
        assert fa[:5] == fb[:5] and fa[6:9] == fb[6:9], (a, b)           # CHROM POS ID REF ALT | FILTER INFO FORMAT
        sa, sb = fa[9].split(":"), fb[9].split(":")
        assert sa[0] == sb[0] and sa[2:] == sb[2:], (a, b)               # GT, DP, AF identical
        if fa[5] != fb[5]:
            nq += 1
            assert abs(float(fa[5]) - float(fb[5])) <= 0.0101            # QUAL: last printed digit
    assert nq <= 0.03 * len(ref)


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
def test_predict_cli_from_pd_text_and_from_reads(tmp_path, golden, golden_weights, small_case, precision):
    from nanosnp_b200 import predict as P
    from nanosnp_b200.dataset import save_reads_npz
    from oracle.pyoracle import write_fasta
    fa = str(tmp_path / "ref.fa")
    write_fasta(fa, {"ctg1": small_case["ref"]})
    cfg = str(P.__file__).replace("predict.py", "config/ont_pileup.yaml")
    # (a) the reference's text seam
    d1 = tmp_path / "pd"; d1.mkdir()
    up = np.char.upper(small_case["ref"].view("S1")).view(np.uint8)
    with open(d1 / "ctg1.pd", "w") as f:
        for p, w in zip(small_case["site_pos"], small_case["windows"]):
            seq = bytes(up[p - 17:p + 16]).decode()
            f.write(" ".join(map(str, w.reshape(-1))) + " \tctg1:%d:%s\t0-\n" % (p, seq))
    out1 = str(tmp_path / "a.vcf")
    P.main(["-config", cfg, "-model_path", str(golden / "ont_pileup_weights.npz"), "-data", str(d1), "-reference", fa, "-output", out1,
            "--precision", precision])
    _compare_vcf(open(out1).read(), (golden / "s2_small.vcf").read_text())
    # (b) packed reads: s1 + s2 on the GPU
    d2 = tmp_path / "reads"; d2.mkdir()
    save_reads_npz(str(d2 / "ctg1.reads.npz"), small_case["reads"], "ctg1", len(small_case["ref"]))
    out2 = str(tmp_path / "b.vcf")
    P.main(["-config", cfg, "-model_path", str(golden / "ont_pileup_weights.npz"), "-data", str(d2), "-reference", fa, "-output", out2,
            "--precision", precision])
    assert open(out2).read() == open(out1).read()
    # (c) a BAM file: BGZF + BAM decode on the host, then the same GPU path
    from nanosnp_b200.bam import write_bam
    bam = str(tmp_path / "reads.bam")
    write_bam(bam, [("ctg1", len(small_case["ref"]))], {"ctg1": small_case["reads"]})
    out3 = str(tmp_path / "c.vcf")
    P.main(["-config", cfg, "-model_path", str(golden / "ont_pileup_weights.npz"), "-data", bam, "-reference", fa, "-output", out3,
            "--precision", precision])
    assert open(out3).read() == open(out1).read()
    # (d) the checkpoint as the reference ships it: torch.save({'encoder': state_dict, 'forward_layer': state_dict}) (utils.py:67-77)
    import torch
    enc, fwd = golden_weights
    ck = str(tmp_path / "ont_pileup.chkpt")
    torch.save({"encoder": {k: torch.from_numpy(np.array(v)) for k, v in enc.items()},
                "forward_layer": {k: torch.from_numpy(np.array(v)) for k, v in fwd.items()}, "epoch": 0}, ck)
    out4 = str(tmp_path / "d.vcf")
    P.main(["-config", cfg, "-model_path", ck, "-data", str(d2), "-reference", fa, "-output", out4, "--precision", precision])
    assert open(out4).read() == open(out1).read()
    with pytest.raises(SystemExit):
        P.main(["-config", cfg, "-model_path", "x", "-data", str(d2), "-reference", fa, "-output", out2, "--no_cuda"])


def test_region_sharded_contig_equals_single_region(tmp_path, golden, golden_weights, small_case):
    """Many small regions (16-bp halos, batches carried across region boundaries) give the same VCF bytes as one region."""
    import io
    from nanosnp_b200.caller import call_contig
    from nanosnp_b200.pipeline import PileupEngine, PileupModelForward, PileupModelWeights
    from nanosnp_b200.runner import RegionRunner
    w = PileupModelWeights(*golden_weights, device="cuda:0")
    runner = RegionRunner(PileupEngine("cuda:0"), PileupModelForward(w, 0))
    outs = []
    for region_len in (1 << 30, 6_500, 1_111):
        sink = io.BytesIO()
        info = call_contig(runner, small_case["reads"], small_case["ref"], "ctg1", sink, 1000, region_len)
        outs.append(sink.getvalue())
        assert info["sites"] == len(small_case["site_pos"])
    assert outs[0] == outs[1] == outs[2]
    ref_lines = [l for l in (golden / "s2_small.vcf").read_text().splitlines() if not l.startswith("#")]
    got = outs[0].decode().splitlines()
    assert len(got) == len(ref_lines) and all(a.split("\t")[:5] == b.split("\t")[:5] for a, b in zip(got, ref_lines))


def test_lstmnetwork_seam(golden_weights, small_case, golden):
    import torch
    from nanosnp_b200.model import LSTMNetwork
    enc, fwd = golden_weights
    net = LSTMNetwork(None, precision="fp32").to("cuda")
    with pytest.raises(RuntimeError):
        net.encoder.load_state_dict({"lstm.weight_ih_l0": enc["lstm.weight_ih_l0"]})
    net.encoder.load_state_dict(enc); net.forward_layer.load_state_dict(fwd)
    net.eval()
    x = torch.from_numpy(small_case["windows"][:1000]).type(torch.FloatTensor).to("cuda")       # predict.py:49
    gt, zy = net.predict(x)
    z = np.load(golden / "s2_small.npz")
    assert gt.shape == (1000, 21) and zy.shape == (1000, 3) and gt.is_cuda
    assert np.abs(gt.detach().cpu().numpy() - z["gt"][:1000]).max() < 2e-5


def test_gpu_site_records_give_identical_vcf_bytes(golden_weights, small_case):
    """site_record_kernel + compact formatter == array formatter on the same GPU probabilities."""
    import io
    import torch
    from nanosnp_b200.caller import call_contig
    from nanosnp_b200.pipeline import PileupEngine, PileupModelForward, PileupModelWeights
    from nanosnp_b200.runner import RegionRunner
    w = PileupModelWeights(*golden_weights, device="cuda:0")
    outs = []
    for records in (False, True):
        runner = RegionRunner(PileupEngine("cuda:0"), PileupModelForward(w, 1), records=records)
        sink = io.BytesIO()
        call_contig(runner, small_case["reads"], small_case["ref"], "ctg1", sink, 1000, 9_000)
        outs.append(sink.getvalue())
    assert outs[0] == outs[1] and len(outs[0]) > 100_000


def test_gpu_text_path_equals_host_text_path(golden_weights, small_case):
    """VCF text assembled on the GPU (vcf_dev.cu, partial batches carried across regions) == host text of the same records."""
    import io
    from nanosnp_b200.caller import call_contig, call_contig_text
    from nanosnp_b200.pipeline import PileupEngine, PileupModelForward, PileupModelWeights
    from nanosnp_b200.runner import RegionRunner
    w = PileupModelWeights(*golden_weights, device="cuda:0")
    runner = RegionRunner(PileupEngine("cuda:0"), PileupModelForward(w, 1), records=True)
    want = io.BytesIO()
    call_contig(runner, small_case["reads"], small_case["ref"], "ctg1", want, 1000, 9_000)
    for region_len in (1 << 30, 9_000, 1_500):
        got = io.BytesIO()
        info = call_contig_text(runner, small_case["reads"], small_case["ref"], "ctg1", got, 1000, region_len)
        assert got.getvalue() == want.getvalue(), region_len
        assert info["sites"] == len(small_case["site_pos"]) and info["vcf_bytes"] == len(want.getvalue())


def test_predict_cli_indexed_bam_and_two_ranks(tmp_path, golden):
    """(a) an indexed BAM is read region by region through the .bai (only the byte range of each region is inflated);
    (b) two ranks (gloo, sharing this GPU) write the same bytes as one process: each formats and pwrite()s its own regions."""
    import os, subprocess, sys
    from nanosnp_b200 import predict as P
    from nanosnp_b200.bam import write_bam
    from nanosnp_b200.synth import SynthConfig, generate_host
    from oracle.pyoracle import write_fasta
    refs, reads = {}, {}
    for i, (name, L) in enumerate([("ctgA", 300_000), ("ctgB", 90_000), ("ctgC", 170_000)]):
        cfg = SynthConfig(contig_len=L, coverage=22.0, contig=name, seed_ref=31 + i, seed_var=41 + i, seed_reads=51 + i, len_median=4000, len_min=300)
        refs[name], reads[name] = generate_host(cfg)
    fa = str(tmp_path / "ref.fa")
    write_fasta(fa, refs)
    cfgp = str(P.__file__).replace("predict.py", "config/ont_pileup.yaml")
    base = ["-config", cfgp, "-model_path", str(golden / "ont_pileup_weights.npz"), "-reference", fa, "--region_len", "50000"]
    plain = str(tmp_path / "plain.bam"); write_bam(plain, [(n, len(r)) for n, r in refs.items()], reads)
    idx = str(tmp_path / "idx.bam"); write_bam(idx, [(n, len(r)) for n, r in refs.items()], reads, index=True)
    o1, o2, o3 = (str(tmp_path / f"{k}.vcf") for k in "abc")
    P.main(base + ["-data", plain, "-output", o1])
    P.main(base + ["-data", idx, "-output", o2])
    a = open(o1, "rb").read()
    assert a == open(o2, "rb").read() and a.count(b"\n") > 10_000
    env = dict(os.environ, NSNP_DIST_BACKEND="gloo", PYTHONPATH=str(golden.parent.parent))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", "-m", "nanosnp_b200.predict"] + base + ["-data", idx, "-output", o3]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(o3, "rb").read() == a
