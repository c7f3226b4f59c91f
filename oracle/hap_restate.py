"""ORACLE (test infrastructure, never on the product path): restatement of NanoSNP's HaplotypeModel s5 stage
(BASELINE configs[4]).

  frequency_features   HaplotypeModel/dataset_dev.py:11-87   (26 statistics x {all, HP1, HP2, unphased} = 104 channels, float64)
  reference_codes      HaplotypeModel/dataset_dev.py:104-118,147-160 (A1 C2 G3 T4, everything else 0)
  HaplotypeModelOracle HaplotypeModel/model_dev.py:59-143    (two 3-layer BiLSTM-256 encoders + Linear, centre rows, dense/tanh, heads)
  predict_rows         HaplotypeModel/predict_dev.py:27-48    (argmax of the 10 genotypes, QUAL -> `ctg\\tpos\\tGT\\tqual`)

Pinned: tests/golden/make_golden_hap.py imports the REAL reference modules from /root/reference (stubs for `tables` /
`ranger21`), runs them on seeded inputs with seeded random-init weights (both shipped checkpoints are missing) and stores
inputs, features, probabilities and the CSV; tests/test_haplotype.py replays them through this file.
"""
from __future__ import annotations

from math import e, log

import numpy as np
import torch
import torch.nn as nn

GT10 = ["AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT"]          # HaplotypeModel/options.py, first ten labels
_CODE = {ord("A"): 1, ord("C"): 2, ord("G"): 3, ord("T"): 4}


def _group_stats(seq, bq, mq):
    """dataset_dev.py:11-55 for one group of read rows: 26 rows of per-position statistics (float64)."""
    cnt = [np.sum(seq == c, axis=0) for c in (1, 2, 3, 4, -1)]
    total = cnt[0] + cnt[1] + cnt[2] + cnt[3] + cnt[4] + 1e-6
    rows = [c / total for c in cnt] + list(cnt)
    for q in (bq, mq):
        sums = [np.sum(q * (seq == c), axis=0) for c in (1, 2, 3, 4)]
        rows += sums + [s / (n + 1e-9) for s, n in zip(sums, cnt[:4])]
    return np.array(rows, dtype=np.float64)


def frequency_features(seq, bq, mq, hp):
    """dataset_dev.py:59-87: [104, L] float64 from read x position int matrices (pad rows are -2 everywhere)."""
    parts = [_group_stats(seq, bq, mq)]
    for tag in (1, 2, 3):
        rows = np.any(hp == tag, axis=1)
        parts.append(_group_stats(seq[rows], bq[rows], mq[rows]) if rows.any() else np.zeros_like(parts[0]))
    return np.concatenate(parts, axis=0)


def reference_codes(ref: np.ndarray, positions1) -> np.ndarray:
    """Reference channel (dataset_dev.py:104-118): code of ref[pos-1]; non-ACGT, lower case and out of range give 0."""
    out = np.zeros(len(positions1), np.int64)
    for k, p in enumerate(positions1):
        if 1 <= p <= len(ref):
            out[k] = _CODE.get(int(ref[p - 1]), 0)
    return out


class _Enc(nn.Module):           # model_dev.py:59-88
    def __init__(self, dim, hidden, layers, dropout):
        super().__init__()
        self.lstm = nn.LSTM(input_size=dim, hidden_size=hidden, num_layers=layers, batch_first=True, dropout=dropout, bidirectional=True)
        self.output_proj = nn.Linear(2 * hidden, hidden, bias=True)

    def forward(self, x):
        out, _ = self.lstm(x)
        return self.output_proj(out)


class _Fwd(nn.Module):           # model_dev.py:90-111
    def __init__(self, hidden, gt_class, zy_class):
        super().__init__()
        self.dense = nn.Linear(2 * hidden, hidden, bias=True)
        self.tanh = nn.Tanh()
        self.genotype_layer = nn.Linear(hidden, gt_class, bias=True)
        self.zygosity_layer = nn.Linear(hidden, zy_class, bias=True)


class HaplotypeModelOracle(nn.Module):
    """Same module tree and construction order as model_dev.LSTMNetwork, so torch.manual_seed(s) gives the same weights."""

    def __init__(self, seed=None, dim=105, hidden=256, layers=3, gt_class=10, zy_class=3, dropout=0.1, pileup_length=33, haplotype_length=11):
        super().__init__()
        if seed is not None:
            torch.manual_seed(int(seed))
        self.pileup_encoder = _Enc(dim, hidden, layers, dropout)
        self.haplotype_encoder = _Enc(dim, hidden, layers, dropout)
        self.forward_layer = _Fwd(hidden, gt_class, zy_class)
        self.pl, self.hl = pileup_length, haplotype_length
        self.eval()

    @torch.no_grad()
    def predict(self, pileup_x, haplotype_x, dtype=torch.float32):          # model_dev.py:133-143
        m = self if dtype == torch.float32 else __import__("copy").deepcopy(self).to(dtype)
        px = torch.as_tensor(pileup_x).to(dtype).permute(0, 2, 1)
        hx = torch.as_tensor(haplotype_x).to(dtype).permute(0, 2, 1)
        a = m.pileup_encoder(px)[:, self.pl // 2, :]
        b = m.haplotype_encoder(hx)[:, self.hl // 2, :]
        out = m.forward_layer.tanh(m.forward_layer.dense(torch.cat((a, b), 1)))
        gt, zy = m.forward_layer.genotype_layer(out), m.forward_layer.zygosity_layer(out)
        return torch.softmax(gt, 1).float(), torch.softmax(zy, 1).float()


def calculate_score(p) -> float:                         # predict_dev.py:22-25 (p is a numpy float32: the arithmetic stays float32 under NumPy 2)
    tmp = max((-10 * log(e, 10)) * log(((1.0 - p) + 1e-300) / (p + 1e-300)) + 10, 0)
    return float(round(tmp, 2))


def predict_rows(contig_pos, gt_prob: np.ndarray) -> str:  # predict_dev.py:39-46
    out = []
    gp = np.max(gt_prob, axis=1); go = np.argmax(gt_prob, axis=1)
    for j, cp in enumerate(contig_pos):
        ctg, pos = cp.split(":")
        out.append(ctg + "\t" + pos + "\t" + GT10[go[j]] + "\t" + str(calculate_score(gp[j])) + "\n")
    return "".join(out)
