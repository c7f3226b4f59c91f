"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's HaplotypeModel s4 read-matrix build,
HaplotypeModel/create_pileup_haplotype.py:23-216, over the pileup columns of oracle/pysam_emul.py.

Pinned: tests/test_hap_groups.py checks it against tests/golden/hapgroups_small.npz, which was produced by the reference's own
code (tests/golden/make_golden_hapgroups.py).  The pileup engine underneath (pysam / htslib) is a restatement: UNPINNED.

subgroup_matrices() handles ONE sub-group (the list of 11-site groups the reference hands to one pysam pileup sweep):
  1. create_pileup_haplotype.py:39-60   groups with a site deeper than max_coverage are removed;
  2. :69-92                             the columns of interest = the hap sites + centre +- flank of every remaining group; the sweep
                                        starts at the first of them (a read that ENDS on that column is not fetched);
  3. :93-131                            one row per QUERY NAME (two alignments of one read share a row, the later one overwrites),
                                        base code A1 C2 G3 T4, -1 for deletion AND reference skip (htslib sets is_del for both),
                                        HP tag (3 = untagged), base quality (0 on a deletion), MAPQ; anything else in SEQ, or a
                                        column deeper than max_coverage, raises inside the reference's bare try/except and the
                                        whole sub-group yields nothing;
  4. :137-205                           per group: rows whose centre is non-zero, ordered by the centre's HP (order inside one HP
                                        class is unspecified: pandas quicksort), for the 11 hap columns and the 2*flank+1 window.
"""
from __future__ import annotations

import numpy as np

BASE = {"A": 1, "C": 2, "G": 3, "T": 4}


def subgroup_matrices(sam, contig, groups, max_coverage=150, flank=16):
    """groups: list of 11 1-based positions each.  Returns a list of dicts (one per surviving group, input order) with
    'positions', 'hap' and 'pile' -> 4 int arrays [depth, 11] / [depth, 2*flank+1] (seq, hp, baseq, mapq); [] when the sub-group dies."""
    groups = [list(map(int, g)) for g in groups]
    sites = sorted({p for g in groups for p in g})
    want = set(sites)
    too_deep = set()
    for col in sam.pileup(contig, sites[0], sites[-1], min_base_quality=0, min_mapping_quality=0):
        p1 = col.pos + 1
        if p1 > sites[-1]:
            break
        if p1 in want and col.n > max_coverage:
            too_deep.add(p1)
    groups = [g for g in groups if not (set(g) & too_deep)]
    if not groups:
        return []
    cols = set()
    for g in groups:
        c = g[len(g) // 2]
        cols.update(range(c - flank, c + flank + 1))
        cols.update(g)
    cols = sorted(cols)
    index = {p: i for i, p in enumerate(cols)}
    rows = {}                                                   # query name -> 4 lists, insertion ordered
    if cols[0] < 0:
        return []                                               # pysam refuses a negative start: caught by the bare except
    for col in sam.pileup(contig, cols[0], cols[-1], min_base_quality=0, min_mapping_quality=0):
        p1 = col.pos + 1
        if p1 > cols[-1]:
            break
        i = index.get(p1)
        if i is None:
            continue
        if col.n > max_coverage:
            return []
        for pr in col.pileups:
            a = pr.alignment
            tag = a.get_tag("HP") if a.has_tag("HP") else 3
            r = rows.setdefault(a.query_name, [[0] * len(cols) for _ in range(4)])
            if not pr.is_del and not pr.is_refskip:
                ch = a.query_sequence[pr.query_position].upper()
                if ch not in BASE:
                    return []
                r[0][i] = BASE[ch]; r[1][i] = tag; r[2][i] = a.query_qualities[pr.query_position]; r[3][i] = a.mapping_quality
            elif pr.is_del:
                r[0][i] = -1; r[1][i] = tag; r[3][i] = a.mapping_quality
    mats = [np.array([r[k] for r in rows.values()], np.int64).reshape(len(rows), len(cols)) for k in range(4)]
    out = []
    for g in groups:
        c = g[len(g) // 2]
        keep = np.nonzero(mats[0][:, index[c]] != 0)[0]
        order = keep[np.argsort(mats[1][keep, index[c]], kind="stable")]
        hap_cols = [index[p] for p in g]
        win_cols = [index[p] for p in range(c - flank, c + flank + 1)]
        out.append({"positions": g,
                    "hap": [m[np.ix_(order, hap_cols)] for m in mats],
                    "pile": [m[np.ix_(order, win_cols)] for m in mats]})
    return out


def canonical_rows(seq, hp, bq, mq, centre_col):
    """Row order inside one HP class is unspecified: sort rows by (centre HP, then content) so two builds can be compared.
    Padding rows (-2) sort last."""
    seq, hp, bq, mq = (np.asarray(a, np.int64) for a in (seq, hp, bq, mq))
    key_hp = np.where(seq[:, centre_col] == -2, 99, hp[:, centre_col])
    keys = [mq[:, j] for j in range(seq.shape[1] - 1, -1, -1)] + [bq[:, j] for j in range(seq.shape[1] - 1, -1, -1)] + \
           [hp[:, j] for j in range(seq.shape[1] - 1, -1, -1)] + [seq[:, j] for j in range(seq.shape[1] - 1, -1, -1)] + [key_hp]
    o = np.lexsort(keys)
    return seq[o], hp[o], bq[o], mq[o]
