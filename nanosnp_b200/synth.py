"""Seeded synthetic reference + ONT-like reads as flat packed arrays.

The arithmetic lives in csrc/synth_core.h and is integer-only, so the host generator (CPU tests) and the
device generator (bench.py at 100 Mb / 30x) produce identical arrays.  This module only builds the
probability tables and lays the reads out (prefix sums of op / base counts).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from statistics import NormalDist

import numpy as np

from . import _lib
from .reads import PackedReads


def _thr(p: float) -> int:
    return int(min(max(p, 0.0), 1.0) * 4294967295.0)


@dataclass
class SynthConfig:
    contig_len: int = 1_000_000
    coverage: float = 30.0
    contig: str = "ctg1"
    seed_ref: int = 11
    seed_var: int = 12
    seed_reads: int = 13
    len_median: float = 8000.0
    len_sigma: float = 0.7
    len_min: int = 500
    len_max: int = 100_000
    sub_rate: float = 0.03
    ins_rate: float = 0.03
    del_rate: float = 0.04
    indel_mean: float = 1.5
    long_indel_rate: float = 0.002       # deliberate handful of > 60 bp indels (tensor_maker.cpp:97)
    snp_rate: float = 1e-3
    lowmapq_frac: float = 0.03
    secondary_frac: float = 0.01
    supp_frac: float = 0.01
    nbase_rate: float = 0.0
    softclip_frac: float = 0.2
    ref_n_period: int = 0
    ref_n_len: int = 0
    ref_lower_period: int = 0
    ref_lower_len: int = 0
    gap_period: int = 0
    gap_len: int = 0
    use_eqx: bool = False
    n_reads: int = 0                      # 0 = derive from coverage
    _tables: dict = field(default_factory=dict, repr=False)

    def tables(self):
        if self._tables:
            return self._tables
        nd = NormalDist()
        qs = np.empty(1025, np.float64)
        for i in range(1025):
            z = nd.inv_cdf(min(max(i / 1024.0, 1e-6), 1 - 1e-6))
            qs[i] = np.exp(np.log(self.len_median) + self.len_sigma * z)
        lq = np.clip(np.round(qs), self.len_min, self.len_max).astype(np.int32)
        p = self.ins_rate + self.del_rate
        j = np.arange(1, 257, dtype=np.float64)
        mcdf = (1.0 - (1.0 - p) ** j) if p > 0 else np.zeros(256)
        mr = np.minimum(np.floor(mcdf * 4294967296.0), 4294967295.0).astype(np.uint32)
        mr[-1] = 4294967295
        r = 1.0 - 1.0 / self.indel_mean if self.indel_mean > 1 else 0.0
        j = np.arange(1, 61, dtype=np.float64)
        ic = np.minimum(np.floor((1.0 - r ** j) * 4294967296.0), 4294967295.0).astype(np.uint32)
        ic[-1] = 4294967295
        self._tables = {"len_quantiles": lq, "mrun_cdf": mr, "indel_cdf": ic}
        return self._tables

    def mean_span(self) -> float:
        lq = self.tables()["len_quantiles"].astype(np.float64)
        return float(((lq[:-1] + lq[1:]) * 0.5).mean())

    def resolved_n_reads(self) -> int:
        if self.n_reads > 0:
            return self.n_reads
        return max(1, int(round(self.coverage * self.contig_len / self.mean_span())))

    def as_struct(self, table_ptrs) -> _lib.SynthCfg:
        c = _lib.SynthCfg()
        c.seed_ref, c.seed_var, c.seed_reads = self.seed_ref, self.seed_var, self.seed_reads
        c.contig_len = self.contig_len
        c.n_reads = self.resolved_n_reads()
        c.sub_thr, c.ins_thr, c.del_thr = _thr(self.sub_rate), _thr(self.ins_rate), _thr(self.del_rate)
        c.snp_thr = _thr(self.snp_rate)
        c.lowmapq_thr, c.secondary_thr, c.supp_thr = _thr(self.lowmapq_frac), _thr(self.secondary_frac), _thr(self.supp_frac)
        c.nbase_thr, c.softclip_thr, c.long_indel_thr = _thr(self.nbase_rate), _thr(self.softclip_frac), _thr(self.long_indel_rate)
        c.len_min = self.len_min
        c.ref_n_period, c.ref_n_len = self.ref_n_period, self.ref_n_len
        c.ref_lower_period, c.ref_lower_len = self.ref_lower_period, self.ref_lower_len
        c.gap_period, c.gap_len = self.gap_period, self.gap_len
        c.use_eqx = int(self.use_eqx)
        c.len_quantiles, c.mrun_cdf, c.indel_cdf = table_ptrs
        return c


def _layout(n_ops: np.ndarray, n_query: np.ndarray):
    cigar_off = np.zeros(n_ops.shape[0] + 1, np.int64)
    np.cumsum(n_ops, out=cigar_off[1:])
    padded = (n_query.astype(np.int64) + 15) // 16 * 16          # every read starts on a 16-base boundary
    seq_off = np.zeros(n_ops.shape[0], np.int64)
    if n_ops.shape[0] > 1:
        np.cumsum(padded[:-1], out=seq_off[1:])
    total = int(padded.sum())
    return cigar_off, seq_off, total


def generate_host(cfg: SynthConfig):
    """Returns (ref uint8[L], PackedReads of numpy arrays).  Plain C loops: use for <= a few Mb."""
    lib = _lib.load()
    t = cfg.tables()
    c = cfg.as_struct([t[k].ctypes.data for k in ("len_quantiles", "mrun_cdf", "indel_cdf")])
    n = int(c.n_reads)
    ref = np.empty(cfg.contig_len, np.uint8)
    _lib.check(lib.nsnp_synth_ref_host(C.byref(c), ref.ctypes.data))
    pos = np.empty(n, np.int32); flag = np.empty(n, np.uint16); mapq = np.empty(n, np.uint8)
    n_ops = np.empty(n, np.int32); n_query = np.empty(n, np.int32)
    _lib.check(lib.nsnp_synth_count_host(C.byref(c), pos.ctypes.data, flag.ctypes.data, mapq.ctypes.data,
                                         n_ops.ctypes.data, n_query.ctypes.data))
    cigar_off, seq_off, total = _layout(n_ops, n_query)
    cigar = np.zeros(int(cigar_off[-1]), np.uint32)
    seq2 = np.zeros(total // 4 + 16, np.uint8)
    nmask = np.zeros(total // 8 + 16, np.uint8) if cfg.nbase_rate > 0 else None
    _lib.check(lib.nsnp_synth_fill_host(C.byref(c), cigar_off.ctypes.data, seq_off.ctypes.data, cigar.ctypes.data,
                                        seq2.ctypes.data, 0 if nmask is None else nmask.ctypes.data))
    return ref, PackedReads(pos, flag, mapq, cigar_off, cigar, seq_off, seq2, nmask)


def generate_device(cfg: SynthConfig, device):
    """Same arrays, generated on the GPU (torch tensors on `device`).  Returns (ref, PackedReads)."""
    import torch
    lib = _lib.load()
    t = cfg.tables()
    dev_tabs = [torch.from_numpy(t[k].view(np.int32)).to(device) for k in ("len_quantiles", "mrun_cdf", "indel_cdf")]
    c = cfg.as_struct([x.data_ptr() for x in dev_tabs])
    n = int(c.n_reads)
    stream = torch.cuda.current_stream(device).cuda_stream
    with torch.cuda.device(device):
        ref = torch.empty(cfg.contig_len, dtype=torch.uint8, device=device)
        _lib.check(lib.nsnp_synth_ref_dev(C.byref(c), ref.data_ptr(), stream))
        pos = torch.empty(n, dtype=torch.int32, device=device)
        flag = torch.empty(n, dtype=torch.int16, device=device)
        mapq = torch.empty(n, dtype=torch.uint8, device=device)
        n_ops = torch.empty(n, dtype=torch.int32, device=device)
        n_query = torch.empty(n, dtype=torch.int32, device=device)
        _lib.check(lib.nsnp_synth_count_dev(C.byref(c), pos.data_ptr(), flag.data_ptr(), mapq.data_ptr(),
                                            n_ops.data_ptr(), n_query.data_ptr(), stream))
        cigar_off = torch.zeros(n + 1, dtype=torch.int64, device=device)
        torch.cumsum(n_ops, 0, out=cigar_off[1:])
        padded = (n_query.to(torch.int64) + 15) // 16 * 16
        seq_off = torch.zeros(n, dtype=torch.int64, device=device)
        if n > 1:
            torch.cumsum(padded[:-1], 0, out=seq_off[1:])
        total = int(padded.sum().item())
        n_cig = int(cigar_off[-1].item())
        cigar = torch.zeros(n_cig, dtype=torch.int32, device=device)
        seq2 = torch.zeros(total // 4 + 16, dtype=torch.uint8, device=device)
        nmask = torch.zeros(total // 8 + 16, dtype=torch.uint8, device=device) if cfg.nbase_rate > 0 else None
        _lib.check(lib.nsnp_synth_fill_dev(C.byref(c), cigar_off.data_ptr(), seq_off.data_ptr(), cigar.data_ptr(),
                                           seq2.data_ptr(), 0 if nmask is None else nmask.data_ptr(), stream))
        torch.cuda.synchronize(device)
    del dev_tabs
    return ref, PackedReads(pos, flag, mapq, cigar_off, cigar, seq_off, seq2, nmask)
