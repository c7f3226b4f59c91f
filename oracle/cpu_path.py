"""ORACLE / CPU BASELINE (never on the product path): the reference's CPU s1+s2 path, stage by stage.

Follows BASELINE.md section 3 / SURVEY.md 8(d) "CPU baseline timing":
  1. mpileup restatement (ours, C; labelled *restated, not samtools*)           make_predict_data.sh:151
  2. the reference's own compiled DNA_CreateCanSnpTensor + DNA_CreatePredictData  make_predict_data.sh:184-215
     (oracle/_ref; falls back to the C restatement oracle/s1_restate.c if the binaries did not travel)
  3. text -> int32 [N,33,18]   (make_bin_predict_data.py:35-55 without PyTables)
  4. predict.py logic on CPU: batches of 1000 through the fp32 model + per-site VCF loop (predict.py:37-195)
Only bench.py (cpu_baseline leg, --impl reference) and tests import this.
"""
from __future__ import annotations

import os
import shutil
import tempfile
import time

import numpy as np

from . import pyoracle as orc
from .s2_restate import PileupModelOracle, predict_vcf, vcf_header


def run_cpu_path(reads, ref: np.ndarray, contig: str, weights, threads: int, batch_size: int = 1000, keep_vcf: bool = False):
    import torch
    torch.set_num_threads(max(1, threads))
    work = tempfile.mkdtemp(prefix="nsnp_cpu_")
    t = {}
    try:
        os.makedirs(work + "/pile")
        mp = f"{work}/pile/{contig}.mpileup"
        t0 = time.perf_counter()
        rows, deepest = orc.mpileup_text(reads, contig, mp)
        t["mpileup_restated"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if orc.have_ref_binaries():
            kind = "reference"
            orc.write_fasta(work + "/ref.fa", {contig: ref})
            _, pd_path = orc.s1_reference(mp, work + "/ref.fa", contig, work, threads=threads)
        else:
            kind = "port"
            pd_path = work + "/" + contig + ".pd"
            orc.s1_restate(mp, contig, ref, want_counts=False, want_windows=False, pd_path=pd_path)
        t["s1_native_tools"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        x, ctgs, pos, refb = orc.parse_pd(pd_path)
        t["text_to_int32"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        model = PileupModelOracle(*weights)
        body = predict_vcf(model, contig, pos, refb, x, batch_size)
        t["predict_py"] = time.perf_counter() - t0
        text = vcf_header([f"{contig}\t{len(ref)}"]) + body if keep_vcf else None
        return {"kind": kind, "n_sites": int(len(pos)), "rows": rows, "stage_s": t, "total_s": float(sum(t.values())),
                "records": body.count("\n"), "vcf": text, "positions": pos}
    finally:
        shutil.rmtree(work, ignore_errors=True)
