"""VCF record text on the GPU (csrc/vcf_dev.cu) for streams of compact site records.

predict.py cuts every contig file into consecutive batches of `batch_size` sites (predict.py:43) and its record logic
reads the first ten sites of a record's batch (predict.py:106,119), so text can only be produced for complete batches:
`GpuVcfText.push` formats the longest batch-aligned prefix of (carried records + new records) and carries the rest on the
device; `flush` formats the last, short batch.  Text and the list of rounding-tie records are fetched with `fetch`, which
also applies the host libc fix-up -- the bytes equal nsnp_vcf_format_contig_records on the same records.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class TextChunk:
    """One formatted chunk: device text + pinned host scalars the kernels write (length, tie count)."""

    def __init__(self, text: torch.Tensor, n_rec: int, ws: torch.Tensor, meta: torch.Tensor, event: torch.cuda.Event):
        self.text, self.n_rec, self.ws, self.meta, self.event = text, n_rec, ws, meta, event


class GpuVcfText:
    def __init__(self, device, contig: str, batch_size: int = 1000, n_buffers: int = 2):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.contig, self.batch = contig, int(batch_size)
        self.cname = contig.encode()
        self.carry: Optional[torch.Tensor] = None            # uint8 [c, 32] on the device, c < batch
        self.n_sites = 0
        self._k = 0
        self._bufs = [dict() for _ in range(n_buffers)]
        self._host = {}

    def _slot(self, n_rec: int):
        b = self._bufs[self._k % len(self._bufs)]
        self._k += 1
        cap = int(self.lib.nsnp_vcf_text_capacity(n_rec, self.cname))
        wsb = int(self.lib.nsnp_vcf_text_workspace_bytes(n_rec))
        if "text" not in b or b["text"].numel() < cap:
            b["text"] = torch.empty(int(cap * 1.2) + 256, dtype=torch.uint8, device=self.device)
        if "ws" not in b or b["ws"].numel() < wsb:
            b["ws"] = torch.empty(int(wsb * 1.2) + 256, dtype=torch.uint8, device=self.device)
        if "meta" not in b:
            b["meta"] = torch.zeros(2, dtype=torch.int64, device=self.device)      # [0] text length
        if "work" not in b or b["work"].shape[0] < n_rec:
            b["work"] = torch.empty((int(n_rec * 1.2) + self.batch, 32), dtype=torch.uint8, device=self.device)
        return b

    def _format(self, rec: torch.Tensor, first_index: int, heads: Optional[torch.Tensor], b) -> TextChunk:
        n = int(rec.shape[0])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nsnp_vcf_text_records(self.cname, rec.data_ptr(), n, 0, first_index, self.batch,
                                                      0 if heads is None else heads.data_ptr(), b["text"].data_ptr(), b["text"].numel(),
                                                      b["meta"].data_ptr(), b["ws"].data_ptr(), b["ws"].numel(), _stream(self.device)))
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return TextChunk(b["text"], n, b["ws"], b["meta"], ev)

    # ---- streaming (single GPU): batch-aligned prefix now, the rest is carried -------------------------------------
    def push(self, rec: torch.Tensor) -> Optional[TextChunk]:
        """rec: uint8 [n, 32] device records of the next region (ascending positions).  Returns the chunk enqueued on the
        current stream, or None when no batch is complete yet."""
        n = int(rec.shape[0])
        if n == 0:
            return None
        c = 0 if self.carry is None else int(self.carry.shape[0])
        m = (c + n) // self.batch * self.batch
        first_index = self.n_sites - c
        self.n_sites += n
        if m == 0:
            self.carry = rec.clone() if self.carry is None else torch.cat([self.carry, rec])
            return None
        b = self._slot(c + n)
        work = b["work"]
        if c:
            work[:c].copy_(self.carry)
        work[c:c + n].copy_(rec)
        chunk = self._format(work[:m], first_index, None, b)
        self.carry = work[m:c + n].clone() if m < c + n else None
        return chunk

    def flush(self) -> Optional[TextChunk]:
        if self.carry is None or int(self.carry.shape[0]) == 0:
            self.carry = None
            return None
        c = int(self.carry.shape[0])
        b = self._slot(c)
        b["work"][:c].copy_(self.carry)
        chunk = self._format(b["work"][:c], self.n_sites - c, None, b)
        self.carry = None
        return chunk

    # ---- arbitrary record ranges with a per-contig table of batch heads (multi-GPU) ------------------------------
    def format_at(self, rec: torch.Tensor, first_index: int, heads: torch.Tensor) -> TextChunk:
        b = self._slot(int(rec.shape[0]))
        return self._format(rec, first_index, heads, b)

    def batch_heads(self, rec: torch.Tensor, first_index: int, heads: torch.Tensor) -> None:
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nsnp_vcf_batch_heads(rec.data_ptr(), int(rec.shape[0]), 0, first_index, self.batch, heads.data_ptr(),
                                                     _stream(self.device)))

    # ---- device -> host ------------------------------------------------------------------------------------------------
    def fetch(self, chunk: TextChunk, stream: Optional[torch.cuda.Stream] = None) -> memoryview:
        """Waits for the chunk, copies exactly its text (and tie list) to pinned host memory and applies the libc tie
        fix-up.  Returns a view of the text bytes, valid until the next fetch."""
        chunk.event.synchronize()
        stream = stream or torch.cuda.current_stream(self.device)
        h = self._host
        if "meta" not in h:
            h["meta"] = torch.zeros(2, dtype=torch.int64).pin_memory()
            h["tiecount"] = torch.zeros(16, dtype=torch.int32).pin_memory()
        cnt_p = C.c_void_p(); ent_p = C.c_void_p(); cap = C.c_int32(0)
        self.lib.nsnp_vcf_text_ties(chunk.ws.data_ptr(), chunk.n_rec, C.byref(cnt_p), C.byref(ent_p), C.byref(cap))
        off_cnt = cnt_p.value - chunk.ws.data_ptr(); off_ent = ent_p.value - chunk.ws.data_ptr()
        with torch.cuda.stream(stream):
            h["meta"].copy_(chunk.meta, non_blocking=True)
            h["tiecount"][:1].copy_(chunk.ws[off_cnt:off_cnt + 4].view(torch.int32), non_blocking=True)
        stream.synchronize()
        n_text = int(h["meta"][0]); n_ties = int(h["tiecount"][0])
        if n_text > chunk.text.numel():
            raise _lib.NsnpError(_lib.E_WORKSPACE, f"VCF text buffer too small: {n_text} bytes")
        if n_ties > cap.value:
            raise _lib.NsnpError(_lib.E_OVERFLOW, f"{n_ties} rounding-tie records in one chunk (capacity {cap.value})")
        if "text" not in h or h["text"].numel() < n_text + 4096:
            h["text"] = torch.empty(int(n_text * 1.3) + 65536, dtype=torch.uint8).pin_memory()
        if n_ties and ("ties" not in h or h["ties"].numel() < n_ties * 64):
            h["ties"] = torch.empty(cap.value * 64, dtype=torch.uint8).pin_memory()
        with torch.cuda.stream(stream):
            h["text"][:n_text].copy_(chunk.text[:n_text], non_blocking=True)
            if n_ties:
                h["ties"][:n_ties * 64].copy_(chunk.ws[off_ent:off_ent + n_ties * 64], non_blocking=True)
        stream.synchronize()
        if n_ties:
            w = self.lib.nsnp_vcf_text_patch_ties(self.cname, h["text"].data_ptr(), n_text, h["text"].numel(), h["ties"].data_ptr(), n_ties)
            if w <= 0 and n_text > 0:
                raise _lib.NsnpError(_lib.E_WORKSPACE, "tie fix-up of the VCF text failed")
            n_text = int(w)
        return memoryview(h["text"].numpy())[:n_text]
