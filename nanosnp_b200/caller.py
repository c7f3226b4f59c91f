"""Whole-contig driver of the GPU s1+s2 path: regions with 16-bp halos -> RegionRunner -> per-contig VCF assembler.

This is what `predict.py` uses for read inputs (BAM / packed reads): it replaces make_predict_data.sh steps 1-5 and the
model + record loop of PileupModel/predict.py for one contig, keeping the reference's 1000-site batch composition.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .predict_io import ContigVcfAssembler
from .reads import PackedReads, max_reference_span, slice_reads
from .runner import RegionRunner
from .shard import plan_regions, read_range_for_region


def _pinned(reads: PackedReads) -> PackedReads:
    return reads.to_torch("cpu", pin=True)


def call_contig(runner: RegionRunner, reads: PackedReads, ref: np.ndarray, contig: str, sink, batch_size: int = 1000,
                region_len: int = 12_500_000, n_threads: int = 0, regions=None) -> dict:
    """reads: numpy PackedReads of one contig (coordinate sorted).  Writes VCF records to `sink` (binary file-like)."""
    L = int(len(ref))
    ref_dev = torch.from_numpy(np.ascontiguousarray(ref)).to(runner.device)
    regions = regions if regions is not None else plan_regions([(contig, L)], region_len)
    span = max_reference_span(reads) + 1
    host_regions = []
    for rg in regions:
        lo, hi = read_range_for_region(reads.pos, span, rg)
        host_regions.append(_pinned(slice_reads(reads, lo, hi)))
    cap = max(4096, max(rg.emit_end - rg.emit_start for rg in regions))

    def pinned_out():
        if runner.records:
            return {"rec": torch.empty((cap, 32), dtype=torch.uint8).pin_memory()}
        return {"pos0": torch.empty(cap, dtype=torch.int32).pin_memory(), "refbase": torch.empty(cap, dtype=torch.uint8).pin_memory(),
                "cov8": torch.empty((cap, 8), dtype=torch.float32).pin_memory(), "gt": torch.empty((cap, 21), dtype=torch.float32).pin_memory(),
                "zy": torch.empty((cap, 3), dtype=torch.float32).pin_memory()}
    host_outs = (pinned_out(), pinned_out())
    asm = ContigVcfAssembler(contig, batch_size, n_threads, sink)

    def consume(k, res):
        if runner.records:
            asm.add_records(res["rec"].numpy())
        else:
            asm.add(res["pos0"].numpy(), res["refbase"].numpy(), res["gt"].numpy(), res["zy"].numpy(), res["cov8"].numpy())
        return None
    n = runner.run_host_many(host_regions, regions, ref_dev, host_outs, consume)
    nbytes = asm.close()
    return {"sites": n, "vcf_bytes": nbytes, "regions": len(regions)}
