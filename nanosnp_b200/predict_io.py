"""VCF text helpers shared by the predict.py drop-in and the whole-contig caller (see predict.py for the CLI)."""
from __future__ import annotations

import os

import numpy as np

from . import _lib


def format_records_into(buf: np.ndarray, contig: str, positions, reference_bases, gt, zy, cov8, batch_size: int, n_threads: int = 0) -> int:
    """Formats all records of consecutive batch_size-site batches into the uint8 array `buf`; returns the byte count
    (raises if buf is too small).  No copies of the inputs when they are already contiguous and typed."""
    lib = _lib.load()
    n = len(positions)
    if n == 0:
        return 0
    pos = np.ascontiguousarray(positions, np.int32); refb = np.ascontiguousarray(reference_bases, np.uint8)
    gt = np.ascontiguousarray(gt, np.float32); zy = np.ascontiguousarray(zy, np.float32); cov8 = np.ascontiguousarray(cov8, np.float32)
    w = lib.nsnp_vcf_format_contig(contig.encode(), n, pos.ctypes.data, refb.ctypes.data, gt.ctypes.data, zy.ctypes.data,
                                   cov8.ctypes.data, batch_size, n_threads or (os.cpu_count() or 1), buf.ctypes.data, buf.shape[0])
    if w < 0:
        raise _lib.NsnpError(_lib.E_WORKSPACE, f"VCF buffer too small: need {-w} bytes")
    return int(w)


def vcf_buffer_bytes(n: int, contig: str) -> int:
    return n * (80 + len(contig)) + 64


def format_records(contig: str, positions, reference_bases, gt: np.ndarray, zy: np.ndarray, cov8: np.ndarray, batch_size: int,
                   n_threads: int = 0) -> bytes:
    """All records of one contig file, consecutive batches of batch_size sites (host arrays) -> VCF text bytes."""
    buf = np.empty(vcf_buffer_bytes(len(positions), contig), np.uint8)
    w = format_records_into(buf, contig, positions, reference_bases, gt, zy, cov8, batch_size, n_threads)
    return buf[:w].tobytes()


class ContigVcfAssembler:
    """Streams one contig's sites region by region into VCF text while keeping the reference's batch composition:
    records are formatted in consecutive batches of `batch_size` sites counted from the contig's first site
    (predict.py:43), so a region boundary in the middle of a batch carries the partial batch over to the next region."""

    def __init__(self, contig: str, batch_size: int = 1000, n_threads: int = 0, sink=None):
        self.contig, self.batch, self.threads, self.sink = contig, batch_size, n_threads, sink
        self.carry = None
        self.n_bytes = 0
        self.n_sites = 0

    def _emit(self, pos1, refb, gt, zy, cov8):
        need = vcf_buffer_bytes(len(pos1), self.contig)
        if getattr(self, "_buf", None) is None or self._buf.shape[0] < need:
            self._buf = np.empty(int(need * 1.1), np.uint8)          # reused across regions
        w = format_records_into(self._buf, self.contig, pos1, refb, gt, zy, cov8, self.batch, self.threads)
        self.n_bytes += w
        if self.sink is not None:
            self.sink.write(self._buf[:w].tobytes())

    def add(self, pos0, refbase, gt, zy, cov8):
        """Host arrays of one region, ascending positions (0-based)."""
        n = len(pos0)
        self.n_sites += n
        pos1 = np.asarray(pos0, np.int32) + 1
        arrs = [pos1, np.asarray(refbase), np.asarray(gt), np.asarray(zy), np.asarray(cov8)]
        start = 0
        if self.carry is not None:
            need = self.batch - len(self.carry[0])
            take = min(need, n)
            merged = [np.concatenate([c, a[:take]]) for c, a in zip(self.carry, arrs)]
            start = take
            if len(merged[0]) == self.batch:
                self._emit(*merged)
                self.carry = None
            else:
                self.carry = merged
                return
        full = (n - start) // self.batch * self.batch
        if full:
            self._emit(*[a[start:start + full] for a in arrs])
        if start + full < n:
            self.carry = [np.array(a[start + full:]) for a in arrs]        # copy: the caller reuses its buffers

    def close(self):
        if self.carry is not None:
            self._emit(*self.carry)
            self.carry = None
        return self.n_bytes


