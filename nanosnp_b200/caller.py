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
from .shard import assign_lpt, gather_to_rank0, plan_regions, read_range_for_region


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


def records_of_regions(runner: RegionRunner, reads: PackedReads, ref: np.ndarray, regions) -> list:
    """Compact site records (uint8 [n, 32] copies, one array per region) of some regions of one contig."""
    assert runner.records, "records_of_regions needs a RegionRunner(records=True)"
    if not regions:
        return []
    ref_dev = torch.from_numpy(np.ascontiguousarray(ref)).to(runner.device)
    span = max_reference_span(reads) + 1
    host_regions = []
    for rg in regions:
        lo, hi = read_range_for_region(reads.pos, span, rg)
        host_regions.append(_pinned(slice_reads(reads, lo, hi)))
    cap = max(4096, max(rg.emit_end - rg.emit_start for rg in regions))
    host_outs = tuple({"rec": torch.empty((cap, 32), dtype=torch.uint8).pin_memory()} for _ in range(2))
    out = [None] * len(regions)

    def consume(k, res):
        out[k] = res["rec"].numpy().copy()            # the pinned buffer is reused two regions later
        return None
    runner.run_host_many(host_regions, regions, ref_dev, host_outs, consume)
    return out


def call_contigs_sharded(contigs, produce, sink, batch_size: int = 1000, region_len: int = 12_500_000, n_threads: int = 0) -> dict:
    """Multi-GPU driver (one process per GPU, torch.distributed already initialised or a single process).

    contigs: [(name, length)] in output order.  Every rank plans the same regions, takes its LPT share and calls
    produce(contig_index, regions_of_that_contig) -> [records per region].  The per-region records are gathered to
    rank 0, put back into (contig, position) order and only then cut into the reference's 1000-site batches per contig
    (predict.py's record logic depends on the batch composition), so the text is identical to a single-GPU run.
    Only rank 0 writes to `sink`; it returns the totals, the other ranks return None."""
    import torch.distributed as dist
    from .predict_io import RECORD_DTYPE
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    regions = plan_regions(contigs, region_len)
    mine = assign_lpt(regions, world)[rank]
    parts = []
    for ci in sorted({regions[i].contig_index for i in mine}):
        rgs = [regions[i] for i in mine if regions[i].contig_index == ci]
        for rec in produce(ci, rgs):
            rec = np.ascontiguousarray(rec)
            view = rec.view(RECORD_DTYPE).reshape(-1) if rec.dtype != RECORD_DTYPE else rec
            parts.append({"contig_index": np.full(len(view), ci, np.int32), "pos": view["pos1"].astype(np.int64), "rec": view})
    merged = gather_to_rank0(parts)
    if rank != 0:
        return None
    sites = nbytes = 0
    if merged:
        for ci, (name, _) in enumerate(contigs):
            sel = merged["rec"][merged["contig_index"] == ci]
            if len(sel) == 0:
                continue
            asm = ContigVcfAssembler(name, batch_size, n_threads, sink)
            asm.add_records(sel)
            nbytes += asm.close()
            sites += len(sel)
    return {"sites": sites, "vcf_bytes": nbytes, "regions": len(regions), "world": world}
