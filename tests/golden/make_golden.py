"""Generates the committed golden fixtures by running the REAL reference (only possible in the build container).

  python tests/golden/make_golden.py

  * s1: the reference's own compiled tools (oracle/_ref, built from /root/reference by oracle/Makefile)
        on SURVEY appendix C-1 / C-2 mpileup text and on a seeded synthetic pileup.
  * s2: the reference's own Python (PileupModel/model.py, predict.py imported from /root/reference with stub
        `ranger` / `tables` modules; PredictDataset replaced by an array-backed dataset because PyTables is
        absent) with the shipped checkpoint PileupModel/models/ont_pileup.chkpt on CPU fp32.
Outputs (tests/golden/): c1.*, c2.*, s1_small.npz, s2_small.npz, s2_small.vcf, s2_tiny_b7.vcf, ont_pileup_weights.npz
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")

from nanosnp_b200.synth import SynthConfig, generate_host   # noqa: E402
from oracle import pyoracle as orc                           # noqa: E402

C1_REF = "CAGATTTTCATATTATGCAGAAAATCTACTTCGCCTGATACGAGTCGGTTATCTTCGGATACTGTATAGTCCCACCTGGT"


def c1_rows():
    rows = []
    for p in range(1, 81):
        R = C1_REF[p - 1]; r = R.lower()
        s = R * 5 + r * 5
        if p == 40: s = "AAACCaaacc"
        elif p == 45: s = "TTTTTtttt+2act+2ac^]TT-3NNNT-3NNN"
        elif p in (46, 47, 48): s = R * 3 + r * 5 + "**"
        elif p == 50: s = R * 5 + r * 4 + "#n"
        rows.append(f"ctg1\t{p}\tN\t10\t{s}\t{'~' * 10}")
    return rows


def c2_ref():
    s = list("ACGT" * 30)
    s[59] = "T"; s[69] = "A"; s[99] = "n"
    return "".join(s)


def c2_rows():
    ref = c2_ref()
    rows = []
    for p in range(1, 121):
        if p == 20:
            continue
        R = ref[p - 1].upper(); r = R.lower()
        if R == "N": R, r = "A", "a"
        s = R * 5 + r * 5
        if p == 60: s = "TTTTTaaaaa"
        elif p == 70: s = "AAAAAttttt"
        elif p == 80: s = "^+" + R * 5 + r * 5
        elif p == 85: s = R * 5 + r * 3 + (R + "+61" + "A" * 61) * 2
        elif p == 100: s = "AAcccccccc"
        rows.append(f"ctg1\t{p}\tN\t10\t{s}\t{'~' * 10}")
    return rows


def run_ref_s1(rows, ref_str, work, **kw):
    os.makedirs(work + "/pile", exist_ok=True)
    mp = work + "/pile/ctg1.mpileup"
    with open(mp, "w") as f:
        f.write("\n".join(rows) + "\n")
    ref = np.frombuffer(ref_str.encode(), np.uint8)
    orc.write_fasta(work + "/ref.fa", {"ctg1": ref})
    tp, pp = orc.s1_reference(mp, work + "/ref.fa", "ctg1", work, **kw)
    return mp, tp, pp


def import_reference_python():
    for name, body in (("ranger", {"Ranger": type("Ranger", (), {})}),
                       ("tables", {"Filters": type("Filters", (), {"__init__": lambda self, *a, **k: None})})):
        m = types.ModuleType(name)
        m.__dict__.update(body)
        sys.modules[name] = m
    sys.path.insert(0, str(REF / "PileupModel"))
    import predict as ref_predict          # noqa
    import model as ref_model              # noqa
    import utils as ref_utils              # noqa
    return ref_predict, ref_model, ref_utils


def main():
    orc.build()
    assert orc.have_ref_binaries(), "oracle/_ref is not built (needs /root/reference)"
    tmp = tempfile.mkdtemp(prefix="golden_")
    # ---- appendix C-1 / C-2 ----
    mp, tp, pp = run_ref_s1(c1_rows(), C1_REF, tmp + "/c1")
    shutil.copy(mp, HERE / "c1.mpileup"); shutil.copy(tp, HERE / "c1.tensor"); shutil.copy(pp, HERE / "c1.pd")
    (HERE / "c1.ref").write_text(C1_REF)
    for tag, kw in (("c2_af012", {}), ("c2_af09", {"snp_min_af": 0.9, "indel_min_af": 0.9})):
        mp, tp, pp = run_ref_s1(c2_rows(), c2_ref(), tmp + "/" + tag, **kw)
        shutil.copy(tp, HERE / f"{tag}.tensor")
    shutil.copy(mp, HERE / "c2.mpileup")
    (HERE / "c2.ref").write_text(c2_ref())

    # ---- seeded synthetic pileup through the reference binaries ----
    cfg = SynthConfig(contig_len=40_000, coverage=18, seed_ref=101, seed_var=102, seed_reads=103, snp_rate=4e-3,
                      nbase_rate=0.002, ref_n_period=9000, ref_n_len=25, ref_lower_period=3100, ref_lower_len=200,
                      gap_period=13_000, gap_len=300, len_median=3000, len_min=200)
    ref, reads = generate_host(cfg)
    work = tmp + "/small"
    os.makedirs(work + "/pile")
    mp = work + "/pile/ctg1.mpileup"
    rows, deepest = orc.mpileup_text(reads, "ctg1", mp)
    orc.write_fasta(work + "/ref.fa", {"ctg1": ref})
    tp, pp = orc.s1_reference(mp, work + "/ref.fa", "ctg1", work)
    x, ctgs, pos, refb = orc.parse_pd(pp)
    print(f"synthetic: {reads.n_reads} reads, {rows} rows, deepest {deepest}, {len(pos)} sites")
    np.savez_compressed(HERE / "s1_small.npz", ref=ref, pos=reads.pos, flag=reads.flag, mapq=reads.mapq,
                        cigar_off=reads.cigar_off, cigar=reads.cigar, seq_off=reads.seq_off, seq2=reads.seq2,
                        nmask=reads.nmask, site_pos=pos.astype(np.int32), site_refbase=refb.astype(np.uint8),
                        windows=x.astype(np.int16), mpileup_rows=np.int64(rows))
    assert np.abs(x).max() < 32767

    # ---- s2 through the reference's own Python ----
    import torch
    import yaml
    torch.set_num_threads(1)
    ref_predict, ref_model, ref_utils = import_reference_python()
    cfgm = ref_utils.AttrDict(yaml.load(open(REF / "PileupModel/config/ont_pileup.yaml"), Loader=yaml.FullLoader))
    net = ref_model.LSTMNetwork(cfgm.model)
    ck = torch.load(REF / "PileupModel/models/ont_pileup.chkpt", map_location="cpu")
    net.encoder.load_state_dict(ck["encoder"]); net.forward_layer.load_state_dict(ck["forward_layer"])
    net.eval()
    w = {"encoder." + k: v.numpy() for k, v in ck["encoder"].items()}
    w.update({"forward_layer." + k: v.numpy() for k, v in ck["forward_layer"].items()})
    np.savez(HERE / "ont_pileup_weights.npz", **w)

    class ArrayDataset(torch.utils.data.Dataset):         # stands in for PredictDataset (dataset.py:118-149)
        def __init__(self, datapath):
            self.x, self.ctgs, self.pos, self.refb = ArrayDataset.payload
        def __getitem__(self, i):
            return self.ctgs[i], self.pos[i], self.refb[i], self.x[i]
        def __len__(self):
            return len(self.x)

    ref_predict.PredictDataset = ArrayDataset
    _orig_loader = ref_predict.DataLoader
    ref_predict.DataLoader = lambda ds, batch_size, shuffle, num_workers: _orig_loader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=0)
    fai = work + "/ref.fa.fai"

    ArrayDataset.payload = (x, ctgs, pos, refb)
    ref_predict.predict(net, ["ctg1.bin"], fai, 1000, str(HERE / "s2_small.vcf"), torch.device("cpu"))
    with torch.no_grad():
        gt, zy = net.predict(torch.from_numpy(x).float())
    np.savez_compressed(HERE / "s2_small.npz", gt=gt.numpy(), zy=zy.numpy())
    # a tiny run with batch_size 7: exercises the IndexError quirk (predict.py:106,119 with < 10 sites per batch)
    k = 61
    ArrayDataset.payload = (x[:k], ctgs[:k], pos[:k], refb[:k])
    ref_predict.predict(net, ["ctg1.bin"], fai, 7, str(HERE / "s2_tiny_b7.vcf"), torch.device("cpu"))
    shutil.copy(fai, HERE / "s2_small.fai")
    print("records:", sum(1 for l in open(HERE / "s2_small.vcf") if not l.startswith("#")),
          sum(1 for l in open(HERE / "s2_tiny_b7.vcf") if not l.startswith("#")))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
