"""HaplotypeModel s4 (BASELINE configs[4]): from the pileup VCF and the HP-tagged BAMs to the
read x position matrices the HaplotypeModel reads.  Drop-in for `HaplotypeModel/make_predict_bins.py` (same CLI):

    H1  select_snp_multiprocess / find_adjacent_sites   select_hetesnp_homosnp.py:122-230   host (NumPy)
    H2  single_group_pileup_haplotype_feature            create_pileup_haplotype.py:23-216   GPU  (csrc/hap_groups.cu)
        multigroups_pileup_haplotype_feature / Run       make_predict_bins.py:75-183         host planning around the kernel
    H3  write_to_bins                                    write_to_bins.py:4-63               host; `.npz` with the HDF5 node names
                                                                                             (+ the PyTables `.bin` when `tables` exists)

The reference sweeps the BAM with pysam twice per sub-group of <= 100 groups; here the alignments of a contig are decoded once
(nanosnp_b200.bam.BamReader(keep_aux=True)), live on the GPU as flat arrays, and one warp per group builds its matrices.  Quirks
that change the output are kept: the last-contig-per-slice bug of find_adjacent_sites (:228), sub-groups that die in the
reference's bare `except` (a non-ACGT letter, a column deeper than max_coverage), rows keyed by query name.  The order of rows
inside one HP class is unspecified in the reference (pandas quicksort); here it is file order.  There is no CPU fallback.
"""
from __future__ import annotations

import argparse
import ctypes as C
import math
import os
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .bam import BamReader, ReadAux
from .pipeline import require_cuda
from .reads import PackedReads

MAJOR_CONTIGS = ["chr" + str(a) for a in list(range(1, 23)) + ["X", "Y"]] + [str(a) for a in list(range(1, 23)) + ["X", "Y"]]
NAMES = ("sequences", "hap", "baseq", "mapq")


# ------------------------------------------------------------------------------------------------ H1
def _read_pileup_vcf(vcf_file: str, quality_threshold: float) -> Dict[str, Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """contig -> (positions ascending, is '0/1', QUAL) of the rows select_hetesnp_homosnp.py:131-150 keeps (first-seen contig order)."""
    rows: Dict[str, dict] = {}
    n_rows = 0
    with open(vcf_file) as f:
        for line in f:
            n_rows += 1
            if line[0] == "#":
                continue
            col = line.strip().split()
            gt = col[9].split(":")[0].replace("|", "/")
            q = float(col[5])
            if gt in ("0/0", "1/1") and q >= quality_threshold:
                continue
            rows.setdefault(col[0], {})[int(col[1])] = (gt == "0/1", q)          # a repeated position: the last row wins (dict)
    if n_rows == 0:
        print("[WARNING] No vcf file found, please check the setting")
    if not rows:
        print("[WARNING] No variant found, please check the setting")
    out = {}
    for ctg, d in rows.items():
        pos = np.fromiter(d.keys(), np.int64, len(d))
        o = np.argsort(pos, kind="stable")
        vals = list(d.values())
        out[ctg] = (pos[o], np.array([v[0] for v in vals], bool)[o], np.array([v[1] for v in vals], np.float64)[o])
    return out


def find_adjacent_sites(pos: np.ndarray, is_het: np.ndarray, qual: np.ndarray, adjacent_size: int, quality_threshold: float,
                        support_quality: float) -> np.ndarray:
    """Groups of one contig, int64 [G, 2*adjacent_size+1]: every row with QUAL < quality_threshold that has adjacent_size
    supporting rows ('0/1', QUAL >= support_quality) on each side, with the nearest ones (select_hetesnp_homosnp.py:186-223)."""
    a = adjacent_size
    sup = np.nonzero(is_het & (qual >= support_quality))[0]
    cand = np.nonzero(qual < quality_threshold)[0]
    nl = np.searchsorted(sup, cand, "left")                     # supporting rows strictly before / after the candidate
    nr = np.searchsorted(sup, cand, "right")
    ok = (nl >= a) & (len(sup) - nr >= a)
    cand, nl, nr = cand[ok], nl[ok], nr[ok]
    if len(cand) == 0:
        return np.zeros((0, 2 * a + 1), np.int64)
    left = sup[nl[:, None] - a + np.arange(a)[None, :]] if a else np.zeros((len(cand), 0), np.int64)
    right = sup[nr[:, None] + np.arange(a)[None, :]] if a else np.zeros((len(cand), 0), np.int64)
    return pos[np.concatenate([left, cand[:, None], right], axis=1)]


def select_snp_multiprocess(vcf_file: str, quality_threshold: float, adjacent_size: int, support_quality: float = 15, nthreads: int = 10,
                            keep_reference_bug: bool = True) -> Dict[str, np.ndarray]:
    """select_hetesnp_homosnp.py:122-181.  The contigs are cut into ceil(n / nthreads)-sized slices and the reference's worker
    returns only the LAST contig of its slice (`adjacent_groups[contig] = ...` sits outside the loop, :228); keep_reference_bug
    reproduces that (with nthreads >= number of contigs nothing is lost)."""
    table = _read_pileup_vcf(vcf_file, quality_threshold)
    order = MAJOR_CONTIGS + list(table.keys())
    contigs = sorted(table.keys(), key=order.index)
    if not contigs:
        return {}
    step = math.ceil(len(contigs) / nthreads)
    out: Dict[str, np.ndarray] = {}
    for dt in range(0, len(contigs), step):
        sl = contigs[dt:dt + step]
        for ctg in (sl[-1:] if keep_reference_bug else sl):
            out[ctg] = find_adjacent_sites(*table[ctg], adjacent_size, quality_threshold, support_quality)
    return out


# ------------------------------------------------------------------------------------------------ planning (make_predict_bins.py)
def plan_subgroups(groups: np.ndarray, step: int = 100, max_gap: int = 1000) -> List[Tuple[int, int]]:
    """[lo, hi) index ranges of the sub-groups one pysam sweep covers (make_predict_bins.py:85-105): up to `step` consecutive
    groups, cut where the next group starts more than max_gap after the previous group's last site."""
    n = len(groups)
    out = []
    dt = 0
    while dt < n:
        i = 1
        while True:
            if dt + i >= n:
                cut = dt + i
            elif int(groups[dt + i][0]) - int(groups[dt + i - 1][-1]) > max_gap:
                cut = dt + i
            elif i == step - 1:
                cut = dt + i + 1
            else:
                i += 1
                continue
            break
        out.append((dt, min(cut, n)))
        dt = cut
    return out


def plan_chunks(n_groups: int, threads: int) -> List[Tuple[int, int]]:
    """make_predict_bins.py:137-141: one output file per ceil(G / threads) groups."""
    step = math.ceil(n_groups / threads)
    return [] if step == 0 else [(dt, min(n_groups, dt + step)) for dt in range(0, n_groups, step)]


# ------------------------------------------------------------------------------------------------ H2 on the GPU
@dataclass
class ContigAlignments:
    """One contig's alignments on the device, ready for nsnp_hap_group_matrices."""
    reads: PackedReads                  # torch, device
    qual: torch.Tensor                  # uint8 [n_bases]
    hp: torch.Tensor                    # uint8 [n]
    end: torch.Tensor                   # int32 [n]
    end_pm: torch.Tensor                # int32 [n] running maximum
    ck: Optional[torch.Tensor]          # int32 [2 * nsnp_hap_checkpoint_count] CIGAR checkpoints
    dup_prev: Optional[torch.Tensor]    # int32 [n] or None
    dup_next: Optional[torch.Tensor]
    device: torch.device

    def struct(self) -> _lib.Reads:
        s = self.reads.as_struct()
        s.qual = self.qual.data_ptr()
        return s


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def upload_alignments(reads: PackedReads, aux: ReadAux, device="cuda:0") -> ContigAlignments:
    lib = _lib.load()
    device = require_cuda(device)
    assert isinstance(reads.pos, np.ndarray)
    n = reads.n_reads
    flag = reads.flag.astype(np.uint32)
    keep = ((flag & 1796) == 0) & ~(((flag & 1) != 0) & ((flag & 2) == 0))        # pysam's stepper "samtools" + ignore_orphans
    prev = nxt = None
    idx = np.nonzero(keep)[0]
    if len(idx):
        h = aux.qhash[idx]
        o = np.argsort(h, kind="stable")                                          # same name together, file order inside
        same = h[o][1:] == h[o][:-1]
        if same.any():
            prev = np.full(n, -1, np.int32); nxt = np.full(n, -1, np.int32)
            a = idx[o][:-1][same]; b = idx[o][1:][same]
            nxt[a] = b; prev[b] = a
    rd = reads.to_torch(device)
    qual = torch.from_numpy(np.ascontiguousarray(aux.qual)).to(device)
    if qual.numel() < reads.n_bases:
        qual = torch.cat([qual, torch.zeros(reads.n_bases - qual.numel(), dtype=torch.uint8, device=device)])
    hp = torch.from_numpy(np.ascontiguousarray(aux.hp)).to(device)
    end, end_pm, ck = read_ends(rd, device)
    return ContigAlignments(rd, qual, hp, end, end_pm, ck, None if prev is None else torch.from_numpy(prev).to(device),
                            None if nxt is None else torch.from_numpy(nxt).to(device), device)


def read_ends(rd: PackedReads, device):
    """(end, running maximum of end, CIGAR checkpoints) of device-resident reads."""
    lib = _lib.load()
    n = rd.n_reads
    end = torch.empty(max(1, n), dtype=torch.int32, device=device)
    ck = torch.empty(2 * int(lib.nsnp_hap_checkpoint_count(n, rd.n_cigar)), dtype=torch.int32, device=device)
    st = rd.as_struct()
    with torch.cuda.device(device):
        _lib.check(lib.nsnp_hap_read_ends(C.byref(st), end.data_ptr(), ck.data_ptr(), _stream(device)))
    return end, (torch.cummax(end[:n], 0).values.contiguous() if n else end), ck


def _launch(al: ContigAlignments, gpos: torch.Tensor, fetch_lo: torch.Tensor, flank: int, cap: int, want_matrices: bool):
    lib = _lib.load()
    G, n_hap = int(gpos.shape[0]), int(gpos.shape[1])
    W = 2 * flank + 1 if flank >= 0 else 0
    dev = al.device
    n_cols = torch.empty((G, n_hap + W), dtype=torch.int32, device=dev)
    depth = torch.empty(G, dtype=torch.int32, device=dev)
    flags = torch.empty(G, dtype=torch.int32, device=dev)
    hap = pile = None
    hp_ptrs = pl_ptrs = None
    if want_matrices:
        hap = [torch.empty((G, cap, n_hap), dtype=torch.int32, device=dev) for _ in range(4)]
        pile = [torch.empty((G, cap, W), dtype=torch.int32, device=dev) for _ in range(4)]
        hp_ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in hap]); pl_ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in pile])
    st = al.struct()
    with torch.cuda.device(dev):
        _lib.check(lib.nsnp_hap_group_matrices(C.byref(st), al.hp.data_ptr(), al.end.data_ptr(), al.end_pm.data_ptr(), 0 if al.ck is None else al.ck.data_ptr(),
                                               0 if al.dup_prev is None else al.dup_prev.data_ptr(), 0 if al.dup_next is None else al.dup_next.data_ptr(),
                                               gpos.data_ptr(), fetch_lo.data_ptr(), G, n_hap, flank, cap, n_cols.data_ptr(), depth.data_ptr(),
                                               flags.data_ptr(), hp_ptrs, pl_ptrs, _stream(dev)))
    return n_cols, depth, flags, hap, pile


@dataclass
class GroupMatrices:
    """Surviving groups of one contig (input order) with their matrices on the device, rows padded with -2 to `cap`."""
    positions: np.ndarray               # int64 [G', n_hap]
    depth: np.ndarray                   # int32 [G'] rows per group
    hap: List[torch.Tensor]             # 4 x int32 [G', cap, n_hap]     sequences, hap, baseq, mapq
    pile: List[torch.Tensor]            # 4 x int32 [G', cap, 2*flank+1]
    source: np.ndarray                  # index of every surviving group in the input


def group_matrices(al: ContigAlignments, groups: np.ndarray, subgroups: List[Tuple[int, int]], max_coverage: int = 150, flank: int = 16) -> GroupMatrices:
    """create_pileup_haplotype.py:23-216 for every sub-group of one contig.  groups: int [G, n_hap] 1-based positions."""
    groups = np.asarray(groups, np.int64)
    G, n_hap = groups.shape
    dev = al.device
    sub = np.zeros(G, np.int64)
    for k, (lo, hi) in enumerate(subgroups):
        sub[lo:hi] = k
    n_sub = len(subgroups)
    # first sweep (:39-60): depth of the group sites; the sweep starts at the sub-group's first site
    first = np.array([groups[lo:hi].min() for lo, hi in subgroups], np.int64)
    gp = torch.from_numpy(groups.astype(np.int32)).to(dev)
    n1, d1, _, _, _ = _launch(al, gp, torch.from_numpy(first[sub].astype(np.int32)).to(dev), -1, 0, False)
    alive = ~(n1.cpu().numpy() > max_coverage).any(axis=1)
    d1 = d1.cpu().numpy()
    idx = np.nonzero(alive)[0]
    empty = GroupMatrices(np.zeros((0, n_hap), np.int64), np.zeros(0, np.int32), [torch.zeros((0, 1, n_hap), dtype=torch.int32, device=dev)] * 4,
                          [torch.zeros((0, 1, 2 * flank + 1), dtype=torch.int32, device=dev)] * 4, np.zeros(0, np.int64))
    if len(idx) == 0:
        return empty
    # second sweep (:69-131) over the surviving groups: it starts at the sub-group's first column of interest
    g2 = groups[idx]; s2 = sub[idx]
    col_lo = np.minimum(g2[:, 0], g2[:, n_hap // 2] - flank)
    start = np.full(n_sub, np.iinfo(np.int64).max, np.int64)
    np.minimum.at(start, s2, col_lo)
    gp2 = torch.from_numpy(g2.astype(np.int32)).to(dev)
    flo2 = torch.from_numpy(start[s2].astype(np.int32)).to(dev)
    cap = max(1, int(d1[idx].max()))              # rows need the centre covered, so they do not depend on where a sweep starts
    n2, depth, flags, hap, pile = _launch(al, gp2, flo2, flank, cap, True)
    n2 = n2.cpu().numpy(); fl = flags.cpu().numpy(); depth = depth.cpu().numpy()
    if (fl & 2).any():
        raise _lib.NsnpError(_lib.E_OVERFLOW, "hap_group_kernel: more rows than the planned capacity")
    # the reference's bare except (:214): a column deeper than max_coverage (assert, :99), a SEQ letter outside ACGT (KeyError, :123)
    # or a negative sweep start (pysam ValueError) kills the WHOLE sub-group
    dead = np.zeros(n_sub, bool)
    bad = ((n2 > max_coverage).any(axis=1)) | ((fl & 1) != 0)
    dead[np.unique(s2[bad])] = True
    dead[start < 0] = True
    keep = np.nonzero(~dead[s2])[0]
    if len(keep) == 0:
        return empty
    kt = torch.from_numpy(keep).to(dev)
    return GroupMatrices(g2[keep], depth[keep].astype(np.int32), [h.index_select(0, kt) for h in hap], [p.index_select(0, kt) for p in pile], idx[keep])


# ------------------------------------------------------------------------------------------------ H3
def chunk_arrays(gm: GroupMatrices, contig: str, lo: int, hi: int, max_pileup_depth: Optional[int] = None,
                 max_haplotype_depth: Optional[int] = None) -> Optional[Dict[str, np.ndarray]]:
    """The arrays write_to_bins.py:4-63 stores for the groups of one chunk (input indices [lo, hi)): groups ordered by candidate
    position, rows padded with -2 to the chunk's deepest group, then cut to max_*_depth."""
    sel = np.nonzero((gm.source >= lo) & (gm.source < hi))[0]
    if len(sel) == 0:
        return None
    n_hap = gm.positions.shape[1]
    cand = gm.positions[sel, n_hap // 2]
    o = sel[np.argsort(cand, kind="stable")]
    d = int(gm.depth[o].max())
    dh = d if max_haplotype_depth is None else min(d, max_haplotype_depth)
    dp = d if max_pileup_depth is None else min(d, max_pileup_depth)
    ot = torch.from_numpy(o).to(gm.hap[0].device)
    out = {}
    for k, name in enumerate(NAMES):
        out["haplotype_" + name] = gm.hap[k].index_select(0, ot)[:, :dh].cpu().numpy()
        out["pileup_" + name] = gm.pile[k].index_select(0, ot)[:, :dp].cpu().numpy()
    out["candidate_positions"] = np.array([[f"{contig}:{p}"] for p in gm.positions[o, n_hap // 2]])
    out["haplotype_positions"] = np.array([[f"{contig}:{p}" for p in row] for row in gm.positions[o]])
    return out


def write_to_bins(output_dir: str, contig: str, arrays: Dict[str, np.ndarray]) -> str:
    """`<ctg>_<first>_<last>.npz` with the node names of write_to_bins.py:45-62 (nanosnp_b200.haplotype.TestDataset reads it), plus the
    PyTables `.bin` itself when `tables` is importable."""
    first = arrays["candidate_positions"][0, 0].split(":")[1]; last = arrays["candidate_positions"][-1, 0].split(":")[1]
    stem = os.path.join(output_dir, f"{contig}_{first}_{last}")
    np.savez(stem + ".npz", **arrays)
    try:
        import tables
    except ImportError:
        return stem + ".npz"
    n_hap = arrays["haplotype_positions"].shape[1]
    with tables.open_file(stem + ".bin", mode="w") as f:
        flt = tables.Filters(complib="blosc:lz4hc", complevel=5)
        for k, v in arrays.items():
            if v.dtype.kind in "US":
                atom = tables.StringAtom(itemsize=30 * (n_hap - 1))
                f.create_earray("/", k, atom=atom, shape=(0,) + v.shape[1:], filters=flt).append(v.astype("S"))
            else:
                f.create_earray("/", k, atom=tables.Atom.from_dtype(np.dtype("int32")), shape=(0,) + v.shape[1:]).append(v)
    return stem + ".npz"


# ------------------------------------------------------------------------------------------------ driver (make_predict_bins.py:121-183)
def load_contig(bam_path: str, contig: str, device="cuda:0", threads: int = 0) -> Optional[ContigAlignments]:
    with BamReader(bam_path, threads, keep_aux=True) as r:
        for _, name, rd in r.contigs({contig}):
            return upload_alignments(rd, r.aux, device)
    return None


def run(pileup_vcf: str, bams: str, output: str, pileup_flanking_size: int = 5, adjacent_size: int = 5, low_quality_threshold: float = 19,
        hete_support_quality: float = 14, max_coverage: int = 150, max_pileup_depth: Optional[int] = None,
        max_haplotype_depth: Optional[int] = None, threads: int = 1, device="cuda:0") -> List[str]:
    """make_predict_bins.Run: `<bams>/<contig>.bam` per contig, one output file per chunk of ceil(G / threads) groups."""
    device = require_cuda(device)
    os.makedirs(output, exist_ok=True)
    groups = select_snp_multiprocess(pileup_vcf, low_quality_threshold, adjacent_size, hete_support_quality, nthreads=threads)
    written = []
    for ctg, g in groups.items():
        chunks = plan_chunks(len(g), threads)
        if not chunks:
            continue
        path = os.path.join(bams, ctg + ".bam")
        assert os.path.exists(path), path
        al = load_contig(path, ctg, device)
        if al is None:
            continue
        subs = [(lo + a, lo + b) for lo, hi in chunks for a, b in plan_subgroups(g[lo:hi])]
        gm = group_matrices(al, g, subs, max_coverage, pileup_flanking_size)
        for lo, hi in chunks:
            arrs = chunk_arrays(gm, ctg, lo, hi, max_pileup_depth, max_haplotype_depth)
            if arrs is None:
                print("multicandidates_pileup_haplotype_feature output is empty")
                continue
            written.append(write_to_bins(output, ctg, arrs))
    print(time.strftime("[%a %b %d %H:%M:%S %Y] Done.", time.localtime()))
    return written


def main(argv=None):
    ap = argparse.ArgumentParser(description="Create pileup and haplotype feature of candidate SNPs for predicting (GPU)")
    ap.add_argument("--pileup_vcf", type=str, required=True)
    ap.add_argument("--bams", type=str, required=True, help="directory of HP-tagged <contig>.bam files")
    ap.add_argument("--output", type=str, required=True)
    ap.add_argument("--pileup_flanking_size", type=int, default=5)
    ap.add_argument("--adjacent_size", type=int, default=5)
    ap.add_argument("--low_quality_threshold", type=int, default=19)
    ap.add_argument("--hete_support_quality", default=14, type=float)
    ap.add_argument("--max_coverage", type=int, default=150)
    ap.add_argument("--max_pileup_depth", type=int, default=None)
    ap.add_argument("--max_haplotype_depth", type=int, default=None)
    ap.add_argument("--threads", "-t", type=int, default=1)
    a = ap.parse_args(argv)
    run(a.pileup_vcf, a.bams, a.output, a.pileup_flanking_size, a.adjacent_size, a.low_quality_threshold, a.hete_support_quality,
        a.max_coverage, a.max_pileup_depth, a.max_haplotype_depth, a.threads)


if __name__ == "__main__":
    main()
