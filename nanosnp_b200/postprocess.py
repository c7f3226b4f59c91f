"""Host-side text stages either side of the third-party phasing step, for file compatibility s1 -> s6:

  select_high_quality_hetesnps   scripts/select_high_quality_hetesnps.py:27-56   per-contig VCFs of confident het SNPs (input of s3)
  merge_pileup_haplotype         scripts/merge.py:15-144                         pileup calls overridden by haplotype-model calls (s6)
  write_pd                       make_predict_data/main.cpp:76-127               the `.pd` text hand-off from GPU arrays

Plain Python over text files: these stages see ~1e5 records per sample and are not on the hot path.  Behaviour (incl. the
quirks listed in the functions) is pinned against the reference scripts by tests/test_postprocess.py.
"""
from __future__ import annotations

import argparse
import os
from typing import Dict, Optional, Tuple

import numpy as np

_P_INFO = '##INFO=<ID=P,Number=0,Type=Flag,Description="Result from pileup model">\n'
_H_INFO = '##INFO=<ID=H,Number=0,Type=Flag,Description="Result from haplotype model">\n'


# ------------------------------------------------------------------------------------------------ s3 input
def select_high_quality_hetesnps(vcf_file: str, out_dir: str, support_quality: float = 15) -> Dict[str, int]:
    """Writes `<out_dir>/<contig>.splited.vcf` with the header and every record whose genotype is neither 0/0 nor 1/1
    ('|' counts as '/') and whose QUAL >= support_quality (select_high_quality_hetesnps.py:27-56).  Header lines are kept
    once each, in first-seen order; contigs without a kept record get no file.  Returns {contig: records}."""
    header, kept = [], {}
    seen = set()
    with open(vcf_file) as fin:
        for row in fin:
            if row.startswith("#"):
                if row not in seen:
                    seen.add(row)
                    header.append(row)
                continue
            col = row.strip().split()
            gt = col[9].split(":")[0].replace("|", "/")
            if gt in ("0/0", "1/1") or float(col[5]) < support_quality:
                continue
            int(col[1])                                   # the reference parses POS: a malformed row raises there too
            kept.setdefault(col[0], []).append(row)
    for ctg, rows in kept.items():
        with open(os.path.join(out_dir, ctg + ".splited.vcf"), "w") as fout:
            fout.writelines(header)
            fout.writelines(rows)
    return {c: len(r) for c, r in kept.items()}


# ------------------------------------------------------------------------------------------------ s6
def _haplotype_call(ref: str, gt: str) -> Optional[Tuple[str, str]]:
    """(ALT, GT) of a haplotype-model genotype string such as 'AC' against REF, or None when merge.py writes nothing
    (merge.py:83-115): hom-ref, and any call whose only alternative allele is an indel symbol."""
    a, b = gt[0], gt[1]
    if ref in gt:
        if a == b:
            return None
        alt, zy = gt.replace(ref, ""), "0/1"
    elif a == b:
        alt, zy = a, "1/1"
    else:
        alt, zy = ",".join(sorted(gt)), "1/2"
    for sym in ("D", "I"):                               # 'D' is tested first: 'ID' loses its D and is written as ALT I
        if sym in alt:
            if zy != "1/2":
                return None
            return gt.replace(sym, ""), "0/1"
    return alt, zy


def merge_pileup_haplotype(pileup_vcf: str, cat_predict: str, output: str, quality: float = 15) -> int:
    """scripts/merge.py:15-144.  Per pileup record:
      QUAL > quality                      kept with INFO=P unless FILTER is RefCall
      QUAL <= quality, no haplotype call  kept with INFO=P when it is not a RefCall and QUAL >= 13
      haplotype QUAL < 13                 same rule as "no haplotype call"
      otherwise                           rewritten from the haplotype genotype with INFO=H (DP / AF of the pileup record), or
                                          dropped (_haplotype_call)
    The two INFO header lines go right after the first header line.  Returns the number of rewritten records."""
    hap: Dict[str, Dict[str, Tuple[str, str]]] = {}
    n_rows = 0
    with open(cat_predict) as fin:
        for row in fin:
            n_rows += 1
            ctg, pos, gt, qual = row.strip().split("\t")
            hap.setdefault(ctg, {})[pos] = (gt, qual)
    if n_rows == 0:
        print("[WARNING] No dp file found, please check the setting")
        print("[WARNING] No dp results found, please check the setting")
    rewritten = 0
    first_header = True
    with open(pileup_vcf) as fin, open(output, "w") as fout:
        for line in fin:
            if line.startswith("#"):
                fout.write(line)
                if first_header:
                    fout.write(_P_INFO + _H_INFO)
                    first_header = False
                continue
            f = line.strip().split("\t")
            ctg, pos, ref, q, filt = f[0], int(f[1]), f[3], float(f[5]), f[6]
            sample = f[-1].split(":")
            depth, af = int(sample[-2]), float(sample[-1])

            def keep_pileup(need_q13: bool):
                if filt != "RefCall" and (q >= 13 or not need_q13):
                    g = list(f)
                    g[7] = "P"
                    fout.write("\t".join(g) + "\n")

            if q > quality:
                keep_pileup(False)
                continue
            call = hap.get(ctg, {}).get(str(pos))
            if call is None or float(call[1]) < 13:
                keep_pileup(True)
                continue
            new = _haplotype_call(ref, call[0])
            if new is None:
                continue
            hq = float(call[1])
            fout.write("%s\t%d\t.\t%s\t%s\t%s\tPASS\tH\tGT:GQ:DP:AF\t%s:%s:%d:%f\n" % (ctg, pos, ref, new[0], str(hq), new[1], str(int(hq)), depth, af))
            rewritten += 1
    return rewritten


# ------------------------------------------------------------------------------------------------ .pd text hand-off
def write_pd(path: str, contig: str, pos1, ref: np.ndarray, windows: np.ndarray) -> int:
    """The reference's `.pd` predict-data text (make_predict_data/main.cpp:89,120-123) from GPU arrays, one line per site:
    `<594 ints, each followed by a space>\\t<contig>:<pos>:<REF33 upper-cased>\\t<depth>-`.
    The third column normally continues with the ALT tallies of `.alt_info`; neither PredictDataset (dataset.py:121-139) nor
    predict.py reads it and the count tensor does not carry inserted sequences, so it ends after `<depth>-`.
    Sites whose centre reference base is not ACGT are dropped as DNA_CreatePredictData does (main.cpp:92)."""
    ref = np.asarray(ref, np.uint8)
    up = np.char.upper(ref.view("S1")).view(np.uint8)
    n = 0
    with open(path, "w") as f:
        for p, w in zip(np.asarray(pos1).tolist(), np.asarray(windows, np.int32)):
            seq = bytes(up[p - 17:p + 16]).decode()
            if len(seq) != 33 or seq[16] not in "ACGT":
                continue
            c = w[16]
            depth = int(c[8] + c[17] - min(c[0], c[1], c[2], c[3]) - min(c[9], c[10], c[11], c[12]))
            f.write(" ".join(map(str, w.reshape(-1).tolist())) + " \t%s:%d:%s\t%d-\n" % (contig, p, seq, depth))
            n += 1
    return n


def main_merge(argv=None):
    ap = argparse.ArgumentParser(description="drop-in for scripts/merge.py")
    ap.add_argument("--pileup_vcf", required=True)
    ap.add_argument("--cat_predict", required=True)
    ap.add_argument("--quality", type=float, default=15)
    ap.add_argument("--output", required=True)
    a = ap.parse_args(argv)
    merge_pileup_haplotype(a.pileup_vcf, a.cat_predict, a.output, a.quality)


def main_select(argv=None):
    ap = argparse.ArgumentParser(description="drop-in for scripts/select_high_quality_hetesnps.py")
    ap.add_argument("--pileup_vcf", required=True)
    ap.add_argument("--support_quality", default=14, type=float)
    ap.add_argument("--output_dir", required=True)
    a = ap.parse_args(argv)
    os.makedirs(a.output_dir, exist_ok=True)
    select_high_quality_hetesnps(a.pileup_vcf, a.output_dir, a.support_quality)
