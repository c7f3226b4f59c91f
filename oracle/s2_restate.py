"""ORACLE (test infrastructure, never on the product path): restatement of NanoSNP's s2 stage.

  PileupModelOracle   PileupModel/model.py:14-39 (BaseEncoder), :56-73 (ForwardLayer), :114-119 (predict)
                      -- a plain PyTorch fp32 CPU module with the checkpoint's own parameter names
                      (PileupModel/utils.py:67-77), evaluated exactly as written (no pruning).
  calculate_score     PileupModel/predict.py:31-34
  vcf_header          PileupModel/predict.py:13-27
  vcf_records         PileupModel/predict.py:45-194 for ONE batch

Pinned: tests/golden/make_golden.py imports the REAL reference modules from /root/reference (with two
stub modules for the absent `ranger` / `tables` packages), runs them on seeded inputs and stores inputs,
probabilities and VCF text under tests/golden/; tests/test_oracle_pinning.py replays those through this file.
"""
from __future__ import annotations

from math import e, log

import numpy as np
import torch
import torch.nn as nn

# PileupModel/options.py:3-6, 8-28, 30
base_idx = {"A": 0, "C": 1, "G": 2, "T": 3}
gt_decoded_labels = ["AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT", "DD", "AD", "CD", "GD", "TD", "II",
                     "AI", "CI", "GI", "TI", "ID"]
zy_decoded_labels = ["0/0", "1/1", "0/1"]


class _Encoder(nn.Module):          # model.py:14-39
    def __init__(self):
        super().__init__()
        self.lstm = nn.LSTM(input_size=18, hidden_size=64, num_layers=2, batch_first=True, dropout=0.3, bidirectional=True)
        self.output_proj = nn.Linear(128, 128, bias=True)

    def forward(self, x):
        out, _ = self.lstm(x)
        return self.output_proj(out)


class _Forward(nn.Module):          # model.py:56-73
    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(128, 256)
        self.genotype_layer = nn.Linear(256, 21)
        self.zygosity_layer = nn.Linear(256, 3)
        self.indel1_layer = nn.Linear(256, 33)
        self.indel2_layer = nn.Linear(256, 33)

    def forward(self, x):
        out = torch.tanh(self.dense(x))[:, 16, :]
        return self.genotype_layer(out), self.zygosity_layer(out)


class PileupModelOracle(nn.Module):
    def __init__(self, encoder_state=None, forward_state=None, seed=None):
        super().__init__()
        if seed is not None:
            torch.manual_seed(seed)
        self.encoder = _Encoder()
        self.forward_layer = _Forward()
        if encoder_state is not None:
            self.encoder.load_state_dict({k: torch.as_tensor(v) for k, v in encoder_state.items()})
        if forward_state is not None:
            self.forward_layer.load_state_dict({k: torch.as_tensor(v) for k, v in forward_state.items()})
        self.eval()

    @torch.no_grad()
    def predict(self, x):            # model.py:114-119
        x = torch.as_tensor(x).to(torch.float32)
        gt, zy = self.forward_layer(self.encoder(x))
        return torch.softmax(gt, 1), torch.softmax(zy, 1)

    @torch.no_grad()
    def predict64(self, x):
        """The same network evaluated in float64 (results rounded to float32): a checker whose value does not depend on
        which fp32 LSTM kernel the host CPU's torch build picks (two GPU boxes differed by 2e-5 in `predict`)."""
        import copy
        m = copy.deepcopy(self).double()
        gt, zy = m.forward_layer(m.encoder(torch.as_tensor(x).to(torch.float64)))
        return torch.softmax(gt, 1).float(), torch.softmax(zy, 1).float()

    def state_dicts(self):
        return ({k: v.detach().clone() for k, v in self.encoder.state_dict().items()},
                {k: v.detach().clone() for k, v in self.forward_layer.state_dict().items()})


def load_weights_npz(path):
    z = np.load(path)
    enc = {k[len("encoder."):]: z[k] for k in z.files if k.startswith("encoder.")}
    fwd = {k[len("forward_layer."):]: z[k] for k in z.files if k.startswith("forward_layer.")}
    return enc, fwd


def calculate_score(probability):    # predict.py:31-34 (probability is a numpy float32 scalar there)
    p = probability
    tmp = max((-10 * log(e, 10)) * log(((1.0 - p) + 1e-300) / (p + 1e-300)) + 10, 0)
    return float(round(tmp, 2))


def vcf_header(fai_lines):           # predict.py:13-27
    out = ["##fileformat=VCFv4.3", '##FILTER=<ID=PASS,Description="All filters passed">',
           '##FILTER=<ID=RefCall,Description="Reference call">']
    for line in fai_lines:
        f = line.strip().split()
        out.append("##contig=<ID={},length={}>".format(f[0], f[1]))
    out += ['##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
            '##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype Quality">',
            '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read Depth">',
            '##FORMAT=<ID=AF,Number=A,Type=Float,Description="Allele Frequency">',
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSample"]
    return "\n".join(out) + "\n"


def vcf_records(ctg_names, positions, reference_bases, position_matrix, gt_output_, zy_output_) -> str:
    """One batch of predict.py:45-194.  position_matrix float32 [n,33,18]; probabilities float32 numpy."""
    out = []
    ctg_names = np.array(ctg_names)
    gt_output_ = np.asarray(gt_output_, np.float32); zy_output_ = np.asarray(zy_output_, np.float32)
    gt_prob = np.max(gt_output_, axis=1); zy_prob = np.max(zy_output_, axis=1)
    gt_output = np.argmax(gt_output_, axis=1); zy_output = np.argmax(zy_output_, axis=1)
    cov_feature = np.asarray(position_matrix, np.float32)[:, 16, [0, 1, 2, 3, 9, 10, 11, 12]]
    fmt = "{0}\t{1}\t.\t{2}\t{3}\t{4}\t{5}\t{6}\t{7}\t{8}\n"
    with np.errstate(all="ignore"):
        for j in range(zy_output.shape[0]):
            try:
                if gt_output[j] >= 10:
                    continue
                contig_name = ctg_names[j]; spos = positions[j]; sref = chr(int(reference_bases[j]))
                alt = gt_decoded_labels[gt_output[j]]; zy = zy_decoded_labels[zy_output[j]]
                cov = cov_feature[j]
                depth = -1 * cov[np.where(cov < 0)].sum()
                support_count = 0
                for base in alt.replace(sref, ""):
                    bidx = base_idx[base]
                    support_count += cov[bidx]
                    support_count += cov[bidx + 4]
                af = support_count / depth
                if af > 1.0:
                    af = 1.0
                gt_qual = calculate_score(gt_prob[j]); zy_qual = calculate_score(zy_prob[j])
                qual = min(gt_qual, zy_qual)
                alt = alt.replace(sref, "")

                def fix_hom():
                    max_ti, max_v = -1, -1
                    for ti in [0, 4, 7, 9]:
                        if gt_decoded_labels[ti][0] == sref:
                            continue
                        if gt_output[ti] > max_v:           # batch array indexed by class index (predict.py:106)
                            max_v = gt_output[ti]; max_ti = ti
                    return gt_decoded_labels[max_ti][0]

                def fix_het():
                    max_ti, max_v = -1, -1
                    for ti in [1, 2, 3, 5, 6, 8]:
                        if gt_output[ti] > max_v:           # predict.py:119
                            max_v = gt_output[ti]; max_ti = ti
                    lab = gt_decoded_labels[max_ti]
                    return lab[1] if lab[0] == sref else lab[0]

                def rec(alt_s, q, flt):
                    return fmt.format(contig_name, spos, sref, alt_s, str(q), flt, ".", "GT:GQ:DP:AF",
                                      zy + ":%s:%d:%f" % (str(int(q)), depth, af))

                if len(alt) == 0:
                    if zy == "0/0":
                        out.append(rec(sref, qual, "RefCall"))
                    elif zy == "1/1":
                        out.append(rec(fix_hom(), zy_qual, "PASS"))
                    elif zy == "0/1":
                        out.append(rec(fix_het(), zy_qual, "PASS"))
                    continue
                elif len(alt) == 1:
                    pass
                else:
                    if alt[0] == alt[1]:
                        alt = alt[0]
                    alt = ",".join(list(alt))
                if len(alt) >= 3 and zy_output[j] != 2:
                    zy = "1/2"
                if alt == sref and zy_output[j] != 0:       # predict.py:143-176 (unreachable: alt never holds sref)
                    if zy == "1/1":
                        out.append(rec(fix_hom(), zy_qual, "PASS"))
                    elif zy == "0/1":
                        out.append(rec(fix_het(), zy_qual, "PASS"))
                    continue
                if alt != sref and zy_output[j] == 0:
                    out.append(rec(alt, gt_qual, "PASS"))
                    continue
                out.append(rec(alt, qual, "PASS"))
            except Exception:      # predict.py:193 bare except
                continue
    return "".join(out)


def predict_vcf(model: PileupModelOracle, contig: str, positions, reference_bases, position_matrix, batch_size=1000) -> str:
    """predict.py:37-195 for one contig file: consecutive batches of `batch_size` sites."""
    chunks = []
    n = len(positions)
    for b in range(0, n, batch_size):
        sl = slice(b, min(n, b + batch_size))
        x = torch.as_tensor(np.asarray(position_matrix[sl])).to(torch.float32)
        gt, zy = model.predict(x)
        chunks.append(vcf_records([contig] * (sl.stop - sl.start), np.asarray(positions[sl]), np.asarray(reference_bases[sl]),
                                  x.numpy(), gt.numpy(), zy.numpy()))
    return "".join(chunks)
