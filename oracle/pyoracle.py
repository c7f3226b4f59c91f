"""ORACLE (test infrastructure): Python front end of the C restatement and of the reference's own binaries.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Nothing here is on the product path.

  mpileup_text()   oracle/mpileup_restate.c   (samtools 1.15.1 restatement -- parity unpinned)
  s1_restate()     oracle/s1_restate.c        (tensor_maker.cpp + main.cpp + make_predict_data, pinned vs _ref)
  s1_reference()   oracle/_ref/DNA_*          (the reference's OWN compiled tools, when present)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
REF_DIR = HERE / "_ref"
REF_BINS = ("DNA_ExtractChrPileupData", "DNA_CreateCanSnpTensor", "DNA_CreatePredictData")


def build() -> None:
    """Compiles liboracle.so and, when /root/reference is present, the reference binaries into oracle/_ref."""
    subprocess.run(["make", "-C", str(HERE), "all"], check=True, stdout=subprocess.DEVNULL)


def have_ref_binaries() -> bool:
    return all((REF_DIR / b).exists() for b in REF_BINS)


class _S1Out(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("flags", C.c_void_p), ("cand_pos", C.c_void_p), ("windows", C.c_void_p),
                ("cand_depth", C.c_void_p), ("cand_cap", C.c_int64)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        _lib = C.CDLL(str(LIB))
        _lib.orc_mpileup_write.restype = C.c_int64
        _lib.orc_mpileup_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_uint32, C.c_int32, C.c_char_p, C.c_int, C.c_void_p]
        _lib.orc_depth_cap.restype = C.c_int64
        _lib.orc_depth_cap.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_int32, C.c_void_p]
        _lib.orc_s1_from_mpileup.restype = C.c_int64
        _lib.orc_s1_from_mpileup.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_double, C.c_double,
                                             C.c_int32, C.c_int32, C.c_void_p, C.c_char_p, C.c_char_p]
    return _lib


def mpileup_text(reads, contig: str, path: str, min_mapq: int = 20, excl_flags: int = 2316, max_depth: int = 144,
                 append: bool = False) -> tuple[int, int]:
    """reads: nanosnp_b200.reads.PackedReads with numpy arrays.  Returns (rows, deepest column).
    max_depth > 0 applies the htslib streaming depth cap; 0 disables it."""
    st = reads.as_struct()
    deepest = C.c_int32(0)
    rows = lib().orc_mpileup_write(C.addressof(st), contig.encode(), min_mapq, excl_flags, max_depth, path.encode(),
                                   int(append), C.addressof(deepest))
    if rows < 0:
        raise OSError(f"orc_mpileup_write failed ({rows})")
    return int(rows), int(deepest.value)


def depth_cap(reads, min_mapq: int = 20, excl_flags: int = 2316, max_depth: int = 144) -> np.ndarray:
    """uint8 [n_reads]: 1 where the pileup engine refuses the read at push time (orc_depth_cap)."""
    st = reads.as_struct()
    out = np.zeros(max(reads.n_reads, 1), np.uint8)
    lib().orc_depth_cap(C.addressof(st), min_mapq, excl_flags, max_depth, out.ctypes.data)
    return out[: reads.n_reads]


@dataclass
class S1Result:
    positions: np.ndarray            # int32 [n], 1-based centres, ascending
    windows: Optional[np.ndarray]    # int32 [n,33,18]
    depth: Optional[np.ndarray]
    counts: Optional[np.ndarray]     # int32 [L,18]
    flags: Optional[np.ndarray]      # uint8 [L]


def s1_restate(mpileup_path: str, contig: str, ref: np.ndarray, snp_min_af=0.12, indel_min_af=0.12, min_coverage=6,
               flank=16, want_counts=True, want_windows=True, tensor_path: Optional[str] = None,
               pd_path: Optional[str] = None, cand_cap: Optional[int] = None) -> S1Result:
    ref = np.ascontiguousarray(ref, np.uint8)
    L = ref.shape[0]
    cap = int(cand_cap if cand_cap is not None else L)
    W = 2 * flank + 1
    counts = np.zeros((L, 18), np.int32) if want_counts else None
    flags = np.zeros(L, np.uint8) if want_counts else None
    pos = np.zeros(cap, np.int32)
    dep = np.zeros(cap, np.int32)
    win = np.zeros((cap, W, 18), np.int32) if want_windows else None
    out = _S1Out(counts.ctypes.data if counts is not None else 0, flags.ctypes.data if flags is not None else 0,
                 pos.ctypes.data, win.ctypes.data if win is not None else 0, dep.ctypes.data, cap)
    n = lib().orc_s1_from_mpileup(mpileup_path.encode(), contig.encode(), ref.ctypes.data, L, snp_min_af, indel_min_af,
                                  min_coverage, flank, C.addressof(out),
                                  tensor_path.encode() if tensor_path else None, pd_path.encode() if pd_path else None)
    if n < 0:
        raise RuntimeError(f"orc_s1_from_mpileup failed ({n})")
    if n > cap:
        raise RuntimeError("candidate capacity too small")
    return S1Result(pos[:n].copy(), None if win is None else win[:n].copy(), dep[:n].copy(), counts, flags)


def write_fasta(path: str, contigs: dict, width: int = 60) -> None:
    """Writes FASTA + .fai for {name: uint8 array} (the reference tools read both: ref_reader.cpp:9-64)."""
    with open(path, "wb") as f, open(path + ".fai", "w") as fai:
        off = 0
        for name, seq in contigs.items():
            hdr = f">{name}\n".encode()
            f.write(hdr)
            off += len(hdr)
            L = len(seq)
            fai.write(f"{name}\t{L}\t{off}\t{width}\t{width + 1}\n")
            b = bytes(seq)
            nfull = L // width
            body = bytearray()
            mv = np.frombuffer(b, np.uint8)
            if nfull:
                lines = np.empty((nfull, width + 1), np.uint8)
                lines[:, :width] = mv[: nfull * width].reshape(nfull, width)
                lines[:, width] = 10
                body += lines.tobytes()
            if L % width:
                body += b[nfull * width:] + b"\n"
            f.write(body)
            off += len(body)


def s1_reference(mpileup_path: str, fasta_path: str, contig: str, workdir: str, snp_min_af=0.12, indel_min_af=0.12,
                 min_coverage=6, flank=16, threads: int = 1) -> tuple[str, str]:
    """Runs the reference's own compiled DNA_CreateCanSnpTensor + DNA_CreatePredictData (make_predict_data.sh:184-215).
    mpileup_path must be `<dir>/<contig>.mpileup`.  Returns (tensor_path, pd_path)."""
    if not have_ref_binaries():
        raise FileNotFoundError("oracle/_ref binaries are not built")
    pile_dir = str(Path(mpileup_path).parent)
    tdir = os.path.join(workdir, "candidate_snp_tensor")
    pdir = os.path.join(workdir, "predict_data")
    os.makedirs(tdir, exist_ok=True)
    subprocess.run([str(REF_DIR / "DNA_CreateCanSnpTensor"), "-reference", fasta_path, "-chr_pileup_dir", pile_dir,
                    "-output_dir", tdir, "-min_af", str(snp_min_af), "-snp_min_af", str(snp_min_af),
                    "-indel_min_af", str(indel_min_af), "-min_coverage", str(min_coverage),
                    "-flanking_base", str(flank), "-num_threads", str(threads), contig],
                   check=True, stderr=subprocess.DEVNULL)
    subprocess.run([str(REF_DIR / "DNA_CreatePredictData"), "-chr_tensor_dir", tdir, "-reference", fasta_path,
                    "-output_dir", pdir, "-num_threads", str(threads), contig], check=True, stderr=subprocess.DEVNULL)
    return os.path.join(tdir, contig + ".tensor"), os.path.join(pdir, contig + ".pd")


def parse_pd(pd_path: str):
    """Text -> arrays, restating make_bin_predict_data.py:35-55 and PredictDataset (dataset.py:118-149) without PyTables."""
    mats, ctgs, poss, refb = [], [], [], []
    with open(pd_path) as f:
        for line in f:
            cols = line.rstrip("\n").split("\t")
            mats.append(np.array(cols[0].split(), dtype=np.int32).reshape(33, 18))
            ctg, p, seq = cols[1].strip().split(":")
            ctgs.append(ctg); poss.append(int(p)); refb.append(ord(seq[16]))
    if not mats:
        return np.zeros((0, 33, 18), np.int32), [], np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.stack(mats), ctgs, np.asarray(poss, np.int64), np.asarray(refb, np.int64)
