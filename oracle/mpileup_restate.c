/*
 * ORACLE (test infrastructure, never on the product path).
 *
 * CPU restatement of `samtools mpileup --min-MQ 20 --min-BQ 0 --reverse-del --excl-flags 2316
 * --max-depth 144` (no -f) as NanoSNP invokes it: dna_sv_tensor/src/scripts/make_predict_data.sh:117,151.
 * samtools/htslib 1.15.1 (Dockerfile:7-8) is a third-party dependency that is NOT vendored under
 * /root/reference and is not installed here, so this file restates its documented output format
 * (SURVEY.md appendix B.1-B.3).  PARITY UNPINNED at this boundary: there is no samtools to compare
 * with and the reference has no tests.  Inside the CIGAR subset of appendix B.4 ([S] M {(I|D) M}* [S],
 * M may be =/X) the format is fully determined by the SAM spec.  Outside it this file follows htslib's
 * published pileup logic as the survey restates it (B.2) and as hand-worked in tests/golden/cigar_cases.txt:
 *   - at the LAST reference column of an op (M/=/X, D or N) the next op is inspected: an insertion run
 *     (consecutive I, with P pads printed as '*' and counted in the length) is reported as +<n><seq>,
 *     followed by -<n>N.. when a deletion comes right after the insertion ("1M2I1D" -> T+2AA-1N);
 *     a deletion run (consecutive D merged) is reported as -<n>N.. unless the current op is itself D;
 *   - leading I/S/H/P before the first reference-consuming op are never reported;
 *   - the --max-depth streaming cap (B.3): a read is not pushed iff the pileup engine is already
 *     assembling the read's own start column (an earlier passing read starts there too) and its node
 *     pool holds more than max_depth nodes; the pool holds every pushed read whose end (exclusive) is
 *     >= that column plus two bookkeeping nodes (list tail + dummy head), i.e. the read is dropped iff
 *     buffered_reads + 2 > max_depth (orc_depth_cap below; max_depth <= 0 disables it).
 *
 * Input: the flat packed read arrays of include/nanosnp_b200.h (host pointers).
 * Output: mpileup text rows `chr \t pos1 \t N \t depth \t bases \t quals`.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_common.h"

typedef struct {
    int64_t read;       /* index into the flat arrays */
    int64_t op;         /* index of the current reference-consuming op (absolute, in cigar[]) */
    int64_t op_end;     /* one past the read's last op */
    int32_t op_off;     /* offset inside the current op */
    int64_t qpos;       /* absolute base index (in seq2) of the current query base */
    int32_t first;      /* 1 until the first column has been printed */
    int32_t rev;
} cursor_t;

static inline int op_of(uint32_t c) { return (int)(c & 15); }
static inline int len_of(uint32_t c) { return (int)(c >> 4); }
static inline int consumes_ref(int op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
static inline int consumes_query(int op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }

static inline char base_char(const nsnp_reads_t* r, int64_t k, int rev) {
    static const char up[4] = {'A', 'C', 'G', 'T'};
    char c;
    if (r->nmask && ((r->nmask[k >> 3] >> (k & 7)) & 1)) c = 'N';
    else c = up[(r->seq2[k >> 2] >> (2 * (k & 3))) & 3];
    return rev ? (char)(c + 32) : c;
}

typedef struct { char* s; size_t n, cap; } sbuf_t;
static void sb_put(sbuf_t* b, char c) {
    if (b->n + 1 >= b->cap) { b->cap = b->cap ? b->cap * 2 : 4096; b->s = (char*)realloc(b->s, b->cap); }
    b->s[b->n++] = c;
}
static void sb_int(sbuf_t* b, int v) { char t[16]; int n = snprintf(t, sizeof t, "%d", v); for (int i = 0; i < n; ++i) sb_put(b, t[i]); }

/* advance `c` to its first reference-consuming op, skipping leading S/I/H/P (never reported: B.2) */
static int seek_first_ref_op(const nsnp_reads_t* r, cursor_t* c) {
    while (c->op < c->op_end) {
        const int op = op_of(r->cigar[c->op]);
        if (consumes_ref(op)) return 1;
        if (consumes_query(op)) c->qpos += len_of(r->cigar[c->op]);
        ++c->op;
    }
    return 0;
}

/*
 * htslib streaming depth cap (appendix B.3).  dropped[i] = 1 for every read the pileup engine refuses at push time.
 * Reads failing the B.1 filter are never pushed and never counted.  Returns the number of dropped reads.
 */
int64_t orc_depth_cap(const nsnp_reads_t* r, int32_t min_mapq, uint32_t excl_flags, int32_t max_depth, uint8_t* dropped)
{
    const int64_t n = r->n_reads;
    memset(dropped, 0, (size_t)n);
    if (max_depth <= 0) return 0;
    int64_t* ends = NULL; size_t n_alive = 0, cap = 0;      /* exclusive ends of the pushed reads still in the pool */
    int64_t n_drop = 0, last_pos = -1;                         /* last_pos: start of the most recent PUSHED read */
    for (int64_t i = 0; i < n; ++i) {
        if ((r->flag[i] & 4) || (r->flag[i] & excl_flags) || r->mapq[i] < min_mapq) continue;
        const int64_t P = r->pos[i];
        int64_t reflen = 0;
        for (int64_t k = r->cigar_off[i]; k < r->cigar_off[i + 1]; ++k)
            if (consumes_ref(op_of(r->cigar[k]))) reflen += len_of(r->cigar[k]);
        const int64_t end = P + (reflen > 0 ? reflen : 1);     /* bam_endpos */
        if (last_pos == P) {
            /* the engine has processed every column < P: nodes with end <= P-1 are back in the pool */
            size_t w = 0;
            for (size_t k = 0; k < n_alive; ++k) if (ends[k] >= P) ends[w++] = ends[k];
            n_alive = w;
            if ((int64_t)n_alive + 2 > max_depth) { dropped[i] = 1; ++n_drop; continue; }
        }
        if (n_alive == cap) { cap = cap ? cap * 2 : 256; ends = (int64_t*)realloc(ends, cap * sizeof *ends); }
        ends[n_alive++] = end;
        last_pos = P;
        if (n_alive > 4096) {                                  /* keep the list short when no same-start read prunes it */
            size_t w = 0;
            for (size_t k = 0; k < n_alive; ++k) if (ends[k] >= P) ends[w++] = ends[k];
            n_alive = w;
        }
    }
    free(ends);
    return n_drop;
}

/*
 * Writes rows for every covered position of the contig to `path` (append = 0 truncates).
 * Returns the number of rows, or -1 on I/O error.  max_depth > 0 applies the streaming cap above.
 */
int64_t orc_mpileup_write(const nsnp_reads_t* r, const char* contig_name, int32_t min_mapq, uint32_t excl_flags,
                          int32_t max_depth, const char* path, int append, int32_t* max_depth_seen)
{
    FILE* out = fopen(path, append ? "a" : "w");
    if (!out) return -1;
    char* iobuf = (char*)malloc(1 << 20);          /* per call: the tests run pieces of a contig on several threads */
    if (iobuf) setvbuf(out, iobuf, _IOFBF, 1 << 20);

    cursor_t* act = NULL; size_t n_act = 0, cap_act = 0;
    sbuf_t bases = {0, 0, 0};
    int64_t next = 0, rows = 0;
    int64_t p = 0;
    int deepest = 0;
    const int64_t n = r->n_reads;
    uint8_t* capped = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
    orc_depth_cap(r, min_mapq, excl_flags, max_depth, capped);
#define DROPPED(i) ((r->flag[i] & 4) || (r->flag[i] & excl_flags) || r->mapq[i] < min_mapq || capped[i])

    while (next < n || n_act > 0) {
        if (n_act == 0) {                     /* jump over uncovered stretches */
            while (next < n && DROPPED(next)) ++next;
            if (next >= n) break;
            p = r->pos[next];
        }
        /* push the reads that start at this column */
        while (next < n && r->pos[next] <= p) {
            if (!DROPPED(next) && r->pos[next] == p) {
                cursor_t c;
                c.read = next; c.op = r->cigar_off[next]; c.op_end = r->cigar_off[next + 1];
                c.op_off = 0; c.qpos = r->seq_off[next]; c.first = 1; c.rev = (r->flag[next] & 16) != 0;
                if (seek_first_ref_op(r, &c)) {
                    if (n_act == cap_act) { cap_act = cap_act ? cap_act * 2 : 256; act = (cursor_t*)realloc(act, cap_act * sizeof *act); }
                    act[n_act++] = c;
                }
            }
            ++next;
        }
        if (n_act == 0) continue;
        if ((int)n_act > deepest) deepest = (int)n_act;

        /* one column */
        bases.n = 0;
        size_t w = 0;
        for (size_t i = 0; i < n_act; ++i) {
            cursor_t c = act[i];
            const uint32_t cg = r->cigar[c.op];
            const int op = op_of(cg), len = len_of(cg);
            if (c.first) {
                int q = r->mapq[c.read]; if (q > 93) q = 93;
                sb_put(&bases, '^'); sb_put(&bases, (char)(q + 33));
                c.first = 0;
            }
            if (op == 2) sb_put(&bases, c.rev ? '#' : '*');                 /* --reverse-del */
            else if (op == 3) sb_put(&bases, c.rev ? '<' : '>');
            else { sb_put(&bases, base_char(r, c.qpos, c.rev)); ++c.qpos; }
            ++c.op_off;
            int done = 0;
            if (c.op_off == len) {
                /* last column of this op: peek at what follows (B.2) */
                int64_t k = c.op + 1;
                const int op2 = k < c.op_end ? op_of(r->cigar[k]) : -1;
                int ins = 0;
                if (op2 == 1) ins = 1;
                else if (op2 == 6 && k + 1 < c.op_end) {                    /* pad first: an insertion only if I ops follow the pads */
                    for (int64_t k2 = k + 1; k2 < c.op_end; ++k2) {
                        const int o = op_of(r->cigar[k2]);
                        if (o == 1) { ins = 1; break; }
                        if (o == 2 || o == 0 || o == 3 || o == 7 || o == 8) break;
                    }
                }
                if (ins) {
                    /* insertion run: consecutive I and P ops; pads print as '*' and count in the length */
                    int tot = 0; int64_t k2 = k;
                    while (k2 < c.op_end && (op_of(r->cigar[k2]) == 1 || op_of(r->cigar[k2]) == 6)) { tot += len_of(r->cigar[k2]); ++k2; }
                    sb_put(&bases, '+'); sb_int(&bases, tot);
                    int64_t q = c.qpos;
                    for (int64_t k3 = k; k3 < k2; ++k3) {
                        const int l3 = len_of(r->cigar[k3]);
                        if (op_of(r->cigar[k3]) == 1) { for (int j = 0; j < l3; ++j) sb_put(&bases, base_char(r, q + j, c.rev)); q += l3; }
                        else for (int j = 0; j < l3; ++j) sb_put(&bases, '*');
                    }
                    if (k2 < c.op_end && op_of(r->cigar[k2]) == 2) {        /* "1M2I1D": the deletion right after the insertion */
                        const int dl = len_of(r->cigar[k2]);
                        sb_put(&bases, '-'); sb_int(&bases, dl);
                        for (int j = 0; j < dl; ++j) sb_put(&bases, c.rev ? 'n' : 'N');
                    }
                } else if (op2 == 2 && op != 2) {
                    int tot = 0; int64_t k2 = k;
                    while (k2 < c.op_end && op_of(r->cigar[k2]) == 2) { tot += len_of(r->cigar[k2]); ++k2; }
                    sb_put(&bases, '-'); sb_int(&bases, tot);
                    for (int j = 0; j < tot; ++j) sb_put(&bases, c.rev ? 'n' : 'N');
                }
                /* move to the next reference-consuming op */
                c.op = k; c.op_off = 0;
                while (c.op < c.op_end && !consumes_ref(op_of(r->cigar[c.op]))) {
                    if (consumes_query(op_of(r->cigar[c.op]))) c.qpos += len_of(r->cigar[c.op]);
                    ++c.op;
                }
                if (c.op >= c.op_end) { sb_put(&bases, '$'); done = 1; }
            }
            if (!done) act[w++] = c;
        }
        fprintf(out, "%s\t%lld\tN\t%d\t", contig_name, (long long)(p + 1), (int)n_act);
        fwrite(bases.s, 1, bases.n, out);
        fputc('\t', out);
        for (size_t i = 0; i < n_act; ++i) fputc('~', out);
        fputc('\n', out);
        ++rows;
        n_act = w;
        ++p;
    }
#undef DROPPED
    free(act); free(bases.s); free(capped);
    const int cerr = fclose(out);
    free(iobuf);
    if (cerr != 0) return -1;
    if (max_depth_seen) *max_depth_seen = deepest;
    return rows;
}
