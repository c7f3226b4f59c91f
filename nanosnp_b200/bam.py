"""BAM front end: BGZF inflate (zlib) + native record decode -> PackedReads per contig.

Replaces the BAM reading that `samtools mpileup` does for the reference (make_predict_data.sh:151): the host decodes the
BAM into the flat packed arrays the GPU path consumes.  A small BGZF/BAM writer is included for tests and for turning
synthetic reads into a file the reference's own tooling could read.
"""
from __future__ import annotations

import ctypes as C
import struct
import zlib
from typing import Dict, List, Tuple

import numpy as np

from . import _lib
from .reads import PackedReads

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_decompress(path: str) -> bytes:
    """All BGZF members of a file, inflated and concatenated (a BGZF file is a series of gzip members)."""
    out = []
    with open(path, "rb") as f:
        data = f.read()
    i, n = 0, len(data)
    while i < n:
        d = zlib.decompressobj(wbits=31)
        out.append(d.decompress(data[i:]))
        used = n - i - len(d.unused_data)
        if used <= 0:
            break
        i += used
    return b"".join(out)


def bgzf_compress(raw: bytes, level: int = 1) -> bytes:
    """Proper BGZF: <= 64 KB blocks, each a gzip member with the BC extra field, plus the EOF marker block."""
    out = []
    for s in range(0, len(raw), 0xFF00):
        chunk = raw[s:s + 0xFF00]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        bsize = len(comp) + 25
        out.append(struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize))
        out.append(comp)
        out.append(struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    out.append(_BGZF_EOF)
    return b"".join(out)


def parse_header(raw: bytes) -> Tuple[str, List[Tuple[str, int]], int]:
    if raw[:4] != b"BAM\x01":
        raise ValueError("not a BAM stream")
    l_text = struct.unpack_from("<i", raw, 4)[0]
    text = raw[8:8 + l_text].decode(errors="replace")
    o = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, o)[0]; o += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", raw, o)[0]; o += 4
        name = raw[o:o + l_name - 1].decode(); o += l_name
        l_ref = struct.unpack_from("<i", raw, o)[0]; o += 4
        refs.append((name, l_ref))
    return text, refs, o


def read_bam(path: str, contigs=None) -> Tuple[List[Tuple[str, int]], Dict[str, PackedReads]]:
    """Returns (reference list, {contig: PackedReads}) for the requested contigs (default: all with reads)."""
    lib = _lib.load()
    raw = bgzf_decompress(path)
    _, refs, first = parse_header(raw)
    buf = np.frombuffer(raw, np.uint8)
    out: Dict[str, PackedReads] = {}
    for rid, (name, _) in enumerate(refs):
        if contigs is not None and name not in contigs:
            continue
        n_cig = C.c_int64(0); n_bases = C.c_int64(0)
        n = lib.nsnp_bam_count(buf.ctypes.data, buf.shape[0], first, rid, C.byref(n_cig), C.byref(n_bases))
        if n < 0:
            raise ValueError(f"{path}: malformed BAM record stream")
        if n == 0:
            continue
        pos = np.empty(n, np.int32); flag = np.empty(n, np.uint16); mapq = np.empty(n, np.uint8)
        cigar_off = np.zeros(n + 1, np.int64); cigar = np.empty(max(1, n_cig.value), np.uint32); seq_off = np.empty(n, np.int64)
        seq2 = np.zeros(n_bases.value // 4 + 16, np.uint8); nmask = np.zeros(n_bases.value // 8 + 16, np.uint8)
        m = lib.nsnp_bam_fill(buf.ctypes.data, buf.shape[0], first, rid, pos.ctypes.data, flag.ctypes.data, mapq.ctypes.data,
                              cigar_off.ctypes.data, cigar.ctypes.data, seq_off.ctypes.data, seq2.ctypes.data, nmask.ctypes.data)
        assert m == n
        if n > 1 and (np.diff(pos) < 0).any():
            raise ValueError(f"{path}: reads of {name} are not coordinate sorted")
        out[name] = PackedReads(pos, flag, mapq, cigar_off, cigar[: n_cig.value], seq_off, seq2, nmask if nmask.any() else None)
    return refs, out


def write_bam(path: str, refs: List[Tuple[str, int]], reads_by_contig: Dict[str, PackedReads], long_cigar_as_tag: int = 65535) -> None:
    """Coordinate-sorted BAM from packed reads (tests / interoperability).  Qualities are written as 0xFF (absent)."""
    code4 = np.array([1, 2, 4, 8], np.uint8)
    parts = [b"BAM\x01"]
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    parts += [struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(refs))]
    for n, l in refs:
        parts += [struct.pack("<i", len(n) + 1), n.encode() + b"\0", struct.pack("<i", l)]
    for rid, (name, _) in enumerate(refs):
        rd = reads_by_contig.get(name)
        if rd is None:
            continue
        for i in range(rd.n_reads):
            cg = rd.cigar[rd.cigar_off[i]:rd.cigar_off[i + 1]].astype(np.uint32)
            ops = cg & 15; lens = (cg >> 4).astype(np.int64)
            l_seq = int(lens[(ops == 0) | (ops == 1) | (ops == 4) | (ops == 7) | (ops == 8)].sum())
            ref_len = int(lens[(ops == 0) | (ops == 2) | (ops == 3) | (ops == 7) | (ops == 8)].sum())
            k = int(rd.seq_off[i]) + np.arange(l_seq)
            c2 = (rd.seq2[k >> 2] >> (2 * (k & 3))) & 3
            c4 = code4[c2]
            if rd.nmask is not None:
                isn = ((rd.nmask[k >> 3] >> (k & 7)) & 1).astype(bool)
                c4 = np.where(isn, 15, c4).astype(np.uint8)
            if l_seq & 1:
                c4 = np.append(c4, 0).astype(np.uint8)
            seq = ((c4[0::2] << 4) | c4[1::2]).astype(np.uint8).tobytes()
            rname = f"r{i}".encode() + b"\0"
            tags = b""
            cig_bytes = cg.tobytes(); n_cig = len(cg)
            if n_cig > long_cigar_as_tag:              # SAM spec: real CIGAR in CG:B,I, placeholder <l_seq>S<ref_len>N in the record
                tags = b"CGBI" + struct.pack("<I", n_cig) + cig_bytes
                cig_bytes = struct.pack("<II", (l_seq << 4) | 4, (ref_len << 4) | 3); n_cig = 2
            body = struct.pack("<iiBBHHHIiii", rid, int(rd.pos[i]), len(rname), int(rd.mapq[i]), 4680, n_cig, int(rd.flag[i]), l_seq, -1, -1, 0)
            body += rname + cig_bytes + seq + b"\xff" * l_seq + tags
            parts += [struct.pack("<i", len(body)), body]
    with open(path, "wb") as f:
        f.write(bgzf_compress(b"".join(parts)))
