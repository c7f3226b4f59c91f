"""Whole-contig driver of the GPU s1+s2 path: regions with 16-bp halos -> RegionRunner -> per-contig VCF assembler.

This is what `predict.py` uses for read inputs (BAM / packed reads): it replaces make_predict_data.sh steps 1-5 and the
model + record loop of PileupModel/predict.py for one contig, keeping the reference's 1000-site batch composition.
"""
from __future__ import annotations

import os
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


def call_contig_text(runner: RegionRunner, reads: PackedReads, ref: np.ndarray, contig: str, sink, batch_size: int = 1000,
                     region_len: int = 12_500_000, regions=None, region_reads=None) -> dict:
    """call_contig with the VCF text assembled on the GPU (csrc/vcf_dev.cu): the host only writes the bytes.
    region_reads: optional ready-made per-region PackedReads (e.g. BamReader.fetch); else `reads` is sliced."""
    from .vcf_text import GpuVcfText
    L = int(len(ref))
    ref_dev = torch.from_numpy(np.ascontiguousarray(ref)).to(runner.device)
    regions = regions if regions is not None else plan_regions([(contig, L)], region_len)
    if region_reads is None:
        span = max_reference_span(reads) + 1
        region_reads = [slice_reads(reads, *read_range_for_region(reads.pos, span, rg)) for rg in regions]
    host_regions = [_pinned(r) for r in region_reads]
    gen = GpuVcfText(runner.device, contig, batch_size)
    nbytes = [0]

    def write(mv):
        nbytes[0] += len(mv)
        if sink is not None:
            sink.write(mv)
    n = runner.run_host_text(host_regions, regions, ref_dev, gen, write)
    return {"sites": n, "vcf_bytes": nbytes[0], "regions": len(regions)}


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


# ---- multi-GPU: records stay on the rank that computed them; only counts, batch heads and text lengths are exchanged ----------
def _all_reduce(t, op):
    """t: CPU tensor.  NCCL needs device tensors: round-trip through the current device."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    if dist.get_backend() == "nccl":
        d = t.cuda()
        dist.all_reduce(d, op=op)
        return d.cpu()
    dist.all_reduce(t, op=op)
    return t


def write_sharded_vcf(out_path: str, header: bytes, contigs, regions, records_by_region: dict, batch_size: int = 1000, device=None) -> dict:
    """Every rank holds the compact site records of ITS regions (records_by_region: region index -> uint8 [n,32] torch tensor
    on `device`, or a host RECORD_DTYPE / uint8 array) and writes their text itself, at the right place of ONE ordered file:

      1. SUM all-reduce of the per-region site counts  -> every region's first site index inside its contig file;
      2. MIN all-reduce of the batch-head table        -> the genotype argmax of the first ten sites of every 1000-site batch of
                                                          every contig (what predict.py's `gt_output[ti]` reads; 10 bytes per batch);
      3. each rank formats its regions (GPU kernels for device records, the host twin otherwise) -- byte-identical to one
         process formatting the merged list, because batches are counted from the contig's first site either way;
      4. SUM all-reduce of the per-region text lengths -> file offsets; every rank pwrite()s its segments, rank 0 the header.

    No record ever crosses ranks (the r01 path pickled all of them to rank 0 and formatted there)."""
    import ctypes as C
    import torch.distributed as dist
    from . import _lib
    from .predict_io import RECORD_DTYPE
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    n_reg = len(regions)
    counts = torch.zeros(n_reg, dtype=torch.int64)
    for i, rec in records_by_region.items():
        counts[i] = int(rec.shape[0])
    counts = _all_reduce(counts, dist.ReduceOp.SUM if world > 1 else None)
    first = [0] * n_reg                                        # first site index of each region inside its contig
    tot = {}
    for i, rg in enumerate(regions):
        first[i] = tot.get(rg.contig_index, 0)
        tot[rg.contig_index] = first[i] + int(counts[i])
    nb = {ci: (n + batch_size - 1) // batch_size for ci, n in tot.items()}
    hoff, o = {}, 0
    for ci in sorted(nb):
        hoff[ci] = o; o += nb[ci]
    on_gpu = any(isinstance(r, torch.Tensor) and r.is_cuda for r in records_by_region.values())
    lib = _lib.load()
    gens = {}
    if on_gpu:
        from .vcf_text import GpuVcfText
        heads = torch.full((max(o, 1), 10), 255, dtype=torch.uint8, device=device)
        for i, rec in records_by_region.items():
            ci = regions[i].contig_index
            g = gens.setdefault(ci, GpuVcfText(device, contigs[ci][0], batch_size))
            if rec.shape[0]:
                g.batch_heads(rec, first[i], heads[hoff[ci]:hoff[ci] + nb[ci]])
        if world > 1:
            if dist.get_backend() == "nccl":
                dist.all_reduce(heads, op=dist.ReduceOp.MIN)
            else:
                heads.copy_(_all_reduce(heads.cpu(), dist.ReduceOp.MIN))
    else:
        heads_np = np.full((max(o, 1), 10), 255, np.uint8)
        for i, rec in records_by_region.items():
            ci = regions[i].contig_index
            r = np.ascontiguousarray(rec).view(RECORD_DTYPE).reshape(-1)
            g = first[i] + np.arange(len(r))
            sel = (g % batch_size) < 10
            heads_np[hoff[ci] + g[sel] // batch_size, g[sel] % batch_size] = r["gt"][sel]
        heads_np = _all_reduce(torch.from_numpy(heads_np), dist.ReduceOp.MIN if world > 1 else None).numpy()
    # 3. text of my regions
    texts = {}
    for i in sorted(records_by_region):
        rec = records_by_region[i]
        ci = regions[i].contig_index
        if rec.shape[0] == 0:
            texts[i] = b""
        elif on_gpu:
            g = gens[ci]
            texts[i] = bytes(g.fetch(g.format_at(rec, first[i], heads[hoff[ci]:hoff[ci] + nb[ci]])))
        else:
            r = np.ascontiguousarray(rec).view(RECORD_DTYPE).reshape(-1)
            cap = len(r) * (96 + len(contigs[ci][0])) + 256
            buf = np.empty(cap, np.uint8)
            h = np.ascontiguousarray(heads_np[hoff[ci]:hoff[ci] + nb[ci]])
            w = lib.nsnp_vcf_format_records_at(contigs[ci][0].encode(), r.ctypes.data, len(r), first[i], batch_size, h.ctypes.data,
                                               buf.ctypes.data, cap)
            if w < 0:
                raise _lib.NsnpError(_lib.E_WORKSPACE, "VCF text buffer too small")
            texts[i] = buf[:w].tobytes()
    # 4. offsets, ordered file
    lens = torch.zeros(n_reg, dtype=torch.int64)
    for i, t in texts.items():
        lens[i] = len(t)
    lens = _all_reduce(lens, dist.ReduceOp.SUM if world > 1 else None)
    offs = np.concatenate([[len(header)], len(header) + np.cumsum(lens.numpy())])
    if rank == 0:
        with open(out_path, "wb") as f:
            f.write(header)
            f.truncate(int(offs[-1]))
    if world > 1:
        dist.barrier()
    fd = os.open(out_path, os.O_WRONLY)
    try:
        for i, t in texts.items():
            if t:
                os.pwrite(fd, t, int(offs[i]))
    finally:
        os.close(fd)
    if world > 1:
        dist.barrier()
    return {"sites": int(counts.sum()), "vcf_bytes": int(offs[-1]) - len(header), "regions": n_reg, "world": world}
