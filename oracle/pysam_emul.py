"""TEST INFRASTRUCTURE ONLY (see oracle/README): a small pure-Python stand-in for the part of pysam that the reference's
HaplotypeModel s4 stage uses (`AlignmentFile.pileup` -> PileupColumn / PileupRead / AlignedSegment), so that the reference's
own create_pileup_haplotype.py can run here (pysam is not installed) and so that oracle/hap_groups_restate.py has a
column iterator.  PARITY UNPINNED for this file: it restates htslib's pileup engine from its published behaviour:

  * stepper "samtools" (pysam's default): reads with flag & (UNMAP 4 | SECONDARY 256 | QCFAIL 512 | DUP 1024) are skipped,
    and paired reads that are not proper pairs (ignore_orphans=True);
  * `pileup(ctg, start, stop)` fetches the reads that overlap the 0-based half-open interval [start, stop) and yields EVERY
    column covered by one of them (truncate=False), reads in file order;
  * bam_plp resolve_cigar2: M/=/X -> query_position; D and N both set is_del = 1, N also is_refskip; query_position is None
    for both in pysam; I/S/H/P consume no reference;
  * `PileupColumn.n` counts every read of the column, deletions and reference skips included.
The depth cap (pysam's max_depth = 8000) is not modelled.
"""
from __future__ import annotations

import bisect
import re
from typing import List, Optional

_CIG = re.compile(r"(\d+)([MIDNSHP=X])")


class Segment:
    """AlignedSegment: what create_pileup_haplotype.py:105-131 reads."""

    def __init__(self, name: str, pos0: int, flag: int, mapq: int, cigar: str, seq: str, qual, hp: Optional[int] = None):
        self.query_name = name
        self.reference_start = int(pos0)
        self.flag = int(flag)
        self.mapping_quality = int(mapq)
        self.cigarstring = cigar
        self.query_sequence = seq
        self.query_qualities = list(qual)
        self._hp = hp
        self.ops = []                       # (ref_start, ref_end, query_start, op)
        x, y = self.reference_start, 0
        for m in _CIG.finditer(cigar):
            l, op = int(m.group(1)), m.group(2)
            if op in "M=X":
                self.ops.append((x, x + l, y, op)); x += l; y += l
            elif op in "DN":
                self.ops.append((x, x + l, y, op)); x += l
            elif op in "IS":
                y += l
        self.reference_end = x
        self._starts = [o[0] for o in self.ops]
        assert y == len(seq), (name, y, len(seq))

    def has_tag(self, t: str) -> bool:
        return t == "HP" and self._hp is not None

    def get_tag(self, t: str):
        if not self.has_tag(t):
            raise KeyError(t)
        return self._hp

    def passes_stepper(self) -> bool:
        if self.flag & (4 | 256 | 512 | 1024):
            return False
        if (self.flag & 1) and not (self.flag & 2):
            return False
        return self.reference_end > self.reference_start

    def at(self, col: int):
        """(is_del, is_refskip, query_position) at the 0-based reference column `col` (must lie inside the alignment)."""
        k = bisect.bisect_right(self._starts, col) - 1
        x0, x1, y0, op = self.ops[k]
        assert x0 <= col < x1
        if op in "M=X":
            return 0, 0, y0 + (col - x0)
        return 1, int(op == "N"), None


class PileupRead:
    def __init__(self, seg: Segment, col: int):
        self.alignment = seg
        self.is_del, self.is_refskip, self.query_position = seg.at(col)


class PileupColumn:
    def __init__(self, col: int, segs: List[Segment]):
        self.pos = col
        self.reference_pos = col
        self._segs = segs
        self.n = len(segs)
        self.nsegments = self.n

    @property
    def pileups(self):
        return [PileupRead(s, self.pos) for s in self._segs]


class AlignmentFile:
    """One contig's reads, coordinate sorted (ties keep the given order)."""
    registry = {}                                       # path -> AlignmentFile (the stub's `pysam.AlignmentFile(path, 'r')`)

    def __init__(self, contig: str, segments: List[Segment]):
        self.contig = contig
        self.segments = list(segments)
        assert all(a.reference_start <= b.reference_start for a, b in zip(self.segments, self.segments[1:]))

    @classmethod
    def open(cls, path, mode="r"):
        return cls.registry[str(path)]

    def fetch_overlapping(self, start: int, stop: int) -> List[Segment]:
        return [s for s in self.segments if s.passes_stepper() and s.reference_start < stop and s.reference_end > start]

    def pileup(self, contig, start, stop, min_base_quality=13, min_mapping_quality=0, **kw):
        assert contig == self.contig
        assert min_base_quality == 0 and min_mapping_quality == 0
        if start < 0 or stop < start:
            raise ValueError("invalid coordinates")
        segs = self.fetch_overlapping(start, stop)
        if not segs:
            return
        active: List[Segment] = []
        nxt = 0
        col = segs[0].reference_start
        last = max(s.reference_end for s in segs)
        while col < last:
            while nxt < len(segs) and segs[nxt].reference_start <= col:
                active.append(segs[nxt]); nxt += 1
            active = [s for s in active if s.reference_end > col]
            if active:
                yield PileupColumn(col, active)
                col += 1
            elif nxt < len(segs):
                col = segs[nxt].reference_start
            else:
                break
