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


RECORD_DTYPE = np.dtype([("gt", np.uint8), ("zy", np.uint8), ("flags", np.uint8), ("ref", np.uint8), ("pos1", np.int32),
                         ("q100_gt", np.int32), ("q100_zy", np.int32), ("depth", np.int32), ("af_q", np.int32),
                         ("p_gt", np.float32), ("p_zy", np.float32)])          # struct nsnp_site_record (32 bytes)
assert RECORD_DTYPE.itemsize == 32
REC_DROP, REC_TIE_GT, REC_TIE_ZY, AF_ONE, AF_NAN, AF_NEG_INF = 1, 2, 4, 1000001, -1, -2


def format_compact_records_into(buf: np.ndarray, contig: str, rec: np.ndarray, batch_size: int, n_threads: int = 0) -> int:
    """Text of compact GPU site records (nsnp_site_records) -> buf; returns the byte count."""
    lib = _lib.load()
    n = len(rec)
    if n == 0:
        return 0
    rec = np.ascontiguousarray(rec)
    assert rec.dtype == RECORD_DTYPE or (rec.dtype == np.uint8 and rec.ndim == 2 and rec.shape[1] == 32)
    w = lib.nsnp_vcf_format_contig_records(contig.encode(), n, rec.ctypes.data, batch_size, n_threads or (os.cpu_count() or 1),
                                           buf.ctypes.data, buf.shape[0])
    if w < 0:
        raise _lib.NsnpError(_lib.E_WORKSPACE, f"VCF buffer too small: need {-w} bytes")
    return int(w)


def records_reference(positions1, reference_bases, gt, zy, cov8) -> np.ndarray:
    """NumPy statement of what site_record_kernel computes (tests; not used by the product path)."""
    from math import log
    n = len(positions1)
    rec = np.zeros(n, RECORD_DTYPE)
    gt = np.asarray(gt, np.float32); zy = np.asarray(zy, np.float32); cov8 = np.asarray(cov8, np.float32)
    rec["gt"] = gt.argmax(1); rec["zy"] = zy.argmax(1); rec["ref"] = reference_bases; rec["pos1"] = positions1
    rec["p_gt"] = gt.max(1); rec["p_zy"] = zy.max(1)
    kscale = -10.0 * (1.0 / log(10.0))
    labels = ["AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT"]
    with np.errstate(all="ignore"):
        for j in range(n):
            cov = cov8[j]
            depth = np.float32(-1.0) * cov[cov < 0].sum(dtype=np.float32)
            rec["depth"][j] = int(depth)
            if rec["gt"][j] >= 10:
                continue
            sref = chr(int(reference_bases[j]))
            support = np.float32(0.0)
            for ch in labels[rec["gt"][j]]:
                if ch != sref:
                    b = "ACGT".index(ch)
                    support = np.float32(support + cov[b]); support = np.float32(support + cov[b + 4])
            af = np.float32(support) / np.float32(depth)
            if af > 1.0:
                rec["af_q"][j] = AF_ONE
            elif af != af:
                rec["af_q"][j] = AF_NAN
            elif np.isneginf(af):
                rec["af_q"][j] = AF_NEG_INF
            else:
                m = abs(float(af)) * 1.0e6
                q = np.floor(m); fr = m - q
                if fr > 0.5 or (fr == 0.5 and q % 2 == 1):
                    q += 1
                rec["af_q"][j] = -int(q) - 3 if np.signbit(af) else int(q)
            flags = 0
            for name, p, tiebit in (("q100_gt", rec["p_gt"][j], REC_TIE_GT), ("q100_zy", rec["p_zy"][j], REC_TIE_ZY)):
                r = np.float32(np.float32(1.0) - p) / np.float32(p)
                x = float(r)
                if not x > 0.0:
                    flags |= REC_DROP
                    continue
                t = kscale * log(x) + 10.0
                t = t if t > 0.0 else 0.0
                h = t * 100.0
                fl = np.floor(h); fr = h - fl
                if abs(fr - 0.5) < 1e-6:
                    flags |= tiebit
                rec[name][j] = int(fl) + (1 if fr > 0.5 else 0)
            rec["flags"][j] = flags
    return rec


class ContigVcfAssembler:
    """Streams one contig's sites region by region into VCF text while keeping the reference's batch composition:
    records are formatted in consecutive batches of `batch_size` sites counted from the contig's first site
    (predict.py:43), so a region boundary in the middle of a batch carries the partial batch over to the next region."""

    def __init__(self, contig: str, batch_size: int = 1000, n_threads: int = 0, sink=None):
        self.contig, self.batch, self.threads, self.sink = contig, batch_size, n_threads, sink
        self.carry = None
        self.n_bytes = 0
        self.n_sites = 0

    def _emit(self, *arrs):
        need = vcf_buffer_bytes(len(arrs[0]), self.contig)
        if getattr(self, "_buf", None) is None or self._buf.shape[0] < need:
            self._buf = np.empty(int(need * 1.1), np.uint8)          # reused across regions
        if len(arrs) == 1:
            w = format_compact_records_into(self._buf, self.contig, arrs[0], self.batch, self.threads)
        else:
            w = format_records_into(self._buf, self.contig, *arrs, self.batch, self.threads)
        self.n_bytes += w
        if self.sink is not None:
            self.sink.write(self._buf[:w].tobytes())

    def add_records(self, rec):
        """Compact GPU records of one region (RECORD_DTYPE or uint8 [n,32]), ascending positions."""
        rec = np.asarray(rec)
        if rec.dtype != RECORD_DTYPE:
            rec = np.ascontiguousarray(rec).view(RECORD_DTYPE).reshape(-1)
        self.n_sites += len(rec)
        self._add_arrays([rec])

    def add(self, pos0, refbase, gt, zy, cov8):
        """Host arrays of one region, ascending positions (0-based)."""
        self.n_sites += len(pos0)
        pos1 = np.asarray(pos0, np.int32) + 1
        self._add_arrays([pos1, np.asarray(refbase), np.asarray(gt), np.asarray(zy), np.asarray(cov8)])

    def _add_arrays(self, arrs):
        n = len(arrs[0])
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


