"""Drop-in for the reference's dataset seam: PileupModel/dataset.py `PredictDataset`.

`PredictDataset(datapath)` yields (contig_name, position, ord(reference_base), int32[33,18]) per site, in file order,
exactly like dataset.py:118-149, from either
  * `<chr>.pd`   -- the reference's own text hand-off (make_predict_data/main.cpp:120-123), or
  * `<chr>.reads.npz` -- flat packed reads (nanosnp_b200.reads.PackedReads fields + `contig`, `contig_len`), in which
                    case the whole s1 stage runs on the GPU (needs `reference=` FASTA and a CUDA device).
PyTables `.bin` files cannot be read here (no PyTables/HDF5 in the image); the error says so.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .reads import FIELDS, PackedReads


def load_fasta(path: str) -> dict:
    """name -> uint8 array (raw case), using the .fai when present (ref_reader.cpp:9-64)."""
    out = {}
    name, chunks = None, []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    out[name] = np.frombuffer(b"".join(chunks), np.uint8).copy()
                name, chunks = line[1:].split()[0].decode(), []
            else:
                chunks.append(line.rstrip(b"\r\n"))
    if name is not None:
        out[name] = np.frombuffer(b"".join(chunks), np.uint8).copy()
    return out


def parse_pd_text(path: str):
    """make_bin_predict_data.py:35-55 + dataset.py:126-135 without the HDF5 detour."""
    mats, ctgs, poss, refb = [], [], [], []
    with open(path) as f:
        for line in f:
            cols = line.rstrip("\n").split("\t")
            if len(cols) < 2:
                continue
            mats.append(np.fromstring(cols[0], dtype=np.int32, sep=" "))
            ctg, p, seq = cols[1].strip().split(":")
            ctgs.append(ctg); poss.append(int(p)); refb.append(ord(seq[16]))
    x = np.stack(mats).reshape(-1, 33, 18) if mats else np.zeros((0, 33, 18), np.int32)
    return x, ctgs, np.asarray(poss, np.int64), np.asarray(refb, np.int64)


def save_reads_npz(path: str, reads: PackedReads, contig: str, contig_len: int) -> None:
    r = reads.to_numpy() if reads.is_torch() else reads
    arrs = {f: getattr(r, f) for f in FIELDS if getattr(r, f) is not None}
    np.savez(path, contig=np.array(contig), contig_len=np.int64(contig_len), **arrs)


def load_reads_npz(path: str):
    z = np.load(path)
    reads = PackedReads(*[z[f] if f in z.files else None for f in FIELDS])
    return reads, str(z["contig"]), int(z["contig_len"])


class PredictDataset:
    def __init__(self, datapath: str, reference: Optional[str] = None, device="cuda:0", engine=None):
        self.x_device = None
        if datapath.endswith(".pd"):
            self.position_matrix, self.contig_names, self.positions, self.reference_bases = parse_pd_text(datapath)
        elif datapath.endswith(".npz"):
            import torch
            from .pipeline import PileupEngine
            if reference is None:
                raise ValueError("packed reads need the reference FASTA (reference=...)")
            reads, contig, contig_len = load_reads_npz(datapath)
            ref = load_fasta(reference)[contig]
            assert len(ref) == contig_len, "reference / reads contig length mismatch"
            eng = engine or PileupEngine(device)
            pos, refbase, x, _, _ = eng.candidate_windows(reads.to_torch(eng.device), torch.from_numpy(ref).to(eng.device))
            self.x_device = x                                    # stays resident for the model
            self.position_matrix = x.cpu().numpy()
            self.positions = pos.cpu().numpy().astype(np.int64) + 1
            self.reference_bases = refbase.cpu().numpy().astype(np.int64)
            self.contig_names = [contig] * len(self.positions)
        elif datapath.endswith(".bin"):
            # predict.py:215 lists `bin_predict_data/<chr>.pd.bin` (PyTables HDF5, make_bin_predict_data.py:90-100).  With
            # PyTables installed the file is read as the reference does; without it the `.pd` text it was made from
            # (make_predict_data.sh step 4 leaves it in ../predict_data/) carries the same arrays.
            try:
                import tables                                    # noqa: F401
            except ImportError:
                tables = None
            if tables is not None:
                with tables.open_file(datapath, "r") as f:       # dataset.py:121-139
                    self.position_matrix = np.asarray(f.root.position_matrix, np.int32)
                    meta = [r[0].decode().split(":") for r in f.root.position]
                self.contig_names = [m[0] for m in meta]
                self.positions = np.asarray([int(m[1]) for m in meta], np.int64)
                self.reference_bases = np.asarray([ord(m[2][16]) for m in meta], np.int64)
            else:
                import os
                base = os.path.basename(datapath)[:-4]
                d = os.path.dirname(os.path.abspath(datapath))
                for cand in (datapath[:-4], os.path.join(d, "..", "predict_data", base)):
                    if cand.endswith(".pd") and os.path.exists(cand):
                        self.position_matrix, self.contig_names, self.positions, self.reference_bases = parse_pd_text(cand)
                        break
                else:
                    raise NotImplementedError("PyTables .bin files need the `tables` package, which is absent, and no `.pd` text of the "
                                              f"same name was found next to {datapath} or in ../predict_data/")
        else:
            raise ValueError(f"unrecognised predict data file: {datapath}")

    def __getitem__(self, i):                                     # dataset.py:141-146
        return self.contig_names[i], self.positions[i], self.reference_bases[i], self.position_matrix[i]

    def __len__(self):
        return len(self.position_matrix)
