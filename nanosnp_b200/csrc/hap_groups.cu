// BASELINE configs[4], s4: the read x position matrices of the HaplotypeModel on the GPU.
//
// Replaces HaplotypeModel/create_pileup_haplotype.py:23-216 (two pysam pileup sweeps + pandas per sub-group): for every 11-site
// group (candidate + 5 heterozygous neighbours each side) one warp finds the alignments that overlap the group's columns, walks
// each alignment's CIGAR once over the sorted columns of interest (11 hap sites + centre +- flank) and writes base code
// (A1 C2 G3 T4, -1 deletion / reference skip, 0 absent), HP tag (3 = untagged), base quality and MAPQ into
// [group][row][column] int32 matrices, rows = alignments whose centre entry is non-zero, ordered by the centre's HP tag
// (file order inside one tag), padded with -2 -- the layout write_to_bins.py:14-30 stores and dataset_dev.py reads.
//
// pysam / htslib behaviour that is reproduced (oracle/pysam_emul.py states it):
//   * stepper "samtools": flag & 1796 skipped, paired-but-not-proper skipped; no MAPQ / base-quality filter;
//   * a sweep `pileup(ctg, start, stop)` only sees alignments with end > start: fetch_lo[g] carries the sub-group's start;
//   * column depth n counts every alignment on the column (deletions and skips included): n_cols[g][j] for the host's
//     max_coverage rules (:46-60 and the assert at :99);
//   * rows are keyed by QUERY NAME (:107-121): alignments of one read (supplementary records) share a row and the later
//     alignment overwrites; dup_prev / dup_next link such alignments (NULL when the contig has none);
//   * a SEQ letter outside ACGT raises KeyError inside the reference's bare try/except (:123): flagged, the host drops the sub-group.
#include "common.cuh"

namespace nsnp {
namespace {

constexpr int kMaxCols = 96;                       // 11 + (2 * 32 + 1) columns at most
constexpr int kWarps = 4;                          // groups per CTA

enum { HG_BAD_BASE = 1, HG_ROW_OVERFLOW = 2 };

__device__ __forceinline__ uint32_t ld_cig(const nsnp_reads_t& rd, int64_t i) {
    return rd.cigar_bits == 16 ? (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(rd.cigar) + i) : __ldg(rd.cigar + i);
}

constexpr int kCkShift = 5;                        // one (reference, query) checkpoint per 32 CIGAR ops (= one warp iteration below)

// checkpoint slots of read r: (cigar_off[r] >> kCkShift) + r + j for block j of its ops
__device__ __forceinline__ int64_t ck_slot0(const nsnp_reads_t& rd, int64_t r) { return (rd.cigar_off[r] >> kCkShift) + r; }

// one warp per read: 32 ops per iteration (coalesced), the block's (reference, query) start is its checkpoint
__global__ void __launch_bounds__(256) read_end_kernel(nsnp_reads_t rd, int32_t* __restrict__ end, int2* __restrict__ ck)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rd.n_reads) return;
    int x = rd.pos[r], y = 0;
    const int64_t c0 = rd.cigar_off[r], c1 = rd.cigar_off[r + 1], s0 = ck_slot0(rd, r);
    for (int64_t k = c0; k < c1; k += 32) {
        if (ck && lane == 0) ck[s0 + ((k - c0) >> kCkShift)] = make_int2(x, y);
        int rl = 0, ql = 0;
        if (k + lane < c1) {
            const uint32_t c = ld_cig(rd, k + lane); const int op = c & 15, len = (int)(c >> 4);
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl = len;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) ql = len;
        }
        x += __reduce_add_sync(0xffffffffu, rl); y += __reduce_add_sync(0xffffffffu, ql);
    }
    if (lane == 0) end[r] = x;
}

__device__ __forceinline__ bool stepper_pass(uint32_t flag) {
    if (flag & 1796u) return false;                 // UNMAP | SECONDARY | QCFAIL | DUP  (pysam stepper "samtools")
    if ((flag & 1u) && !(flag & 2u)) return false;  // ignore_orphans
    return true;
}

struct Cell { int v, hp, bq, mq; };

// Walks alignment r over the sorted 0-based columns tg[0..nt) and calls f(column index, cell) for every column it covers.
// A column that lies far ahead is reached through the read's checkpoints (binary search, then <= 32 ops) instead of op by op.
// Returns true when it met a SEQ letter outside ACGT on one of the columns.
template <class F>
__device__ bool walk_read(const nsnp_reads_t& rd, const int2* __restrict__ ck, const uint8_t* __restrict__ qual, int tag, int64_t r,
                          const int* tg, int nt, F&& f)
{
    int x = rd.pos[r], y = 0, ti = 0;
    while (ti < nt && tg[ti] < x) ++ti;
    if (ti >= nt) return false;
    const int mq = rd.mapq[r];
    const int64_t so = rd.seq_off[r];
    const int64_t k0 = rd.cigar_off[r], ke = rd.cigar_off[r + 1], s0 = ck_slot0(rd, r);
    bool bad = false;
    int sought = -1;
    for (int64_t k = k0; k < ke && ti < nt; ++k) {
        if (ck && ti != sought && tg[ti] - x >= 128) {                 // seek: the last checkpoint at or before the column
            sought = ti;
            int64_t lo = (k - k0) >> kCkShift, hi = (ke - 1 - k0) >> kCkShift;
            while (lo < hi) { const int64_t m = (lo + hi + 1) >> 1; if (ck[s0 + m].x <= tg[ti]) lo = m; else hi = m - 1; }
            if (k0 + (lo << kCkShift) > k) { const int2 c = ck[s0 + lo]; k = k0 + (lo << kCkShift); x = c.x; y = c.y; }
        }
        const uint32_t c = ld_cig(rd, k); const int op = c & 15, len = (int)(c >> 4);
        if (op == 0 || op == 7 || op == 8) {
            while (ti < nt && tg[ti] < x + len) {
                const int64_t b = so + y + (tg[ti] - x);
                const int code = (rd.seq2[b >> 2] >> (2 * (b & 3))) & 3;
                if (rd.nmask && ((rd.nmask[b >> 3] >> (b & 7)) & 1)) bad = true;
                f(ti, Cell{code + 1, tag, qual ? (int)qual[b] : 0, mq});
                ++ti;
            }
            x += len; y += len;
        } else if (op == 2 || op == 3) {
            while (ti < nt && tg[ti] < x + len) { f(ti, Cell{-1, tag, 0, mq}); ++ti; }
            x += len;
        } else if (op == 1 || op == 4) {
            y += len;
        }
    }
    return bad;
}

struct GroupArgs {
    nsnp_reads_t rd;
    const uint8_t* hp;            // [n_reads] HP tag, 0 = none
    const int32_t* end;           // [n_reads] exclusive reference end
    const int32_t* end_pm;        // [n_reads] running maximum of end
    const int2* ck;               // CIGAR checkpoints (nsnp_hap_read_ends), may be NULL
    const int32_t* dup_prev;      // [n_reads] previous / next alignment of the same query name (file order), -1; may be NULL
    const int32_t* dup_next;
    const int32_t* gpos;          // [G][n_hap] 1-based, ascending
    const int32_t* fetch_lo;      // [G] start of the pysam sweep that covers the group
    int64_t n_groups;
    int n_hap, flank, cap;        // flank < 0: hap columns only (first sweep); cap = rows per group in the outputs
    int32_t* n_cols;              // [G][n_hap + 2*flank+1] column depth
    int32_t* depth;               // [G] rows
    int32_t* gflags;              // [G]
    int32_t* hap[4];              // [G][cap][n_hap]      seq, hp, baseq, mapq   (NULL: counts only)
    int32_t* pile[4];             // [G][cap][2*flank+1]
};

__global__ void __launch_bounds__(kWarps * 32) hap_group_kernel(const GroupArgs a)
{
    __shared__ int s_tg[kWarps][kMaxCols];        // merged sorted 0-based columns
    __shared__ int8_t s_hc[kWarps][kMaxCols];     // column -> hap column / window column, -1 = none
    __shared__ int8_t s_wc[kWarps][kMaxCols];
    __shared__ int s_n[kWarps][kMaxCols];         // depth per merged column
    __shared__ int s_nt[kWarps];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * kWarps + w;
    if (g >= a.n_groups) return;
    const nsnp_reads_t& rd = a.rd;
    int* tg = s_tg[w]; int8_t* hc = s_hc[w]; int8_t* wc = s_wc[w]; int* ncol = s_n[w];
    const int32_t* gp = a.gpos + g * a.n_hap;
    const int centre = gp[a.n_hap / 2] - 1;                       // 0-based
    const int W = a.flank >= 0 ? 2 * a.flank + 1 : 0;
    if (lane == 0) {                                              // merge the two sorted column lists (tiny)
        int i = 0, j = 0, n = 0;
        while (i < a.n_hap || j < W) {
            const int ph = i < a.n_hap ? gp[i] - 1 : INT32_MAX, pw = j < W ? centre - a.flank + j : INT32_MAX;
            const int p = ph < pw ? ph : pw;
            tg[n] = p; hc[n] = -1; wc[n] = -1;
            if (ph == p) hc[n] = (int8_t)i++;
            if (pw == p) wc[n] = (int8_t)j++;
            ++n;
        }
        s_nt[w] = n;
    }
    __syncwarp();
    const int nt = s_nt[w];
    for (int i = lane; i < nt; i += 32) ncol[i] = 0;
    __syncwarp();
    const int lo_col = tg[0], hi_col = tg[nt - 1];
    // alignments that can cover a column: pos <= hi_col and (running max of end) > lo_col
    int64_t lo, hi;
    {
        int64_t l = 0, h = rd.n_reads;
        while (l < h) { const int64_t m = (l + h) >> 1; if (rd.pos[m] <= hi_col) l = m + 1; else h = m; }
        hi = l;
        l = 0; h = rd.n_reads;
        while (l < h) { const int64_t m = (l + h) >> 1; if (a.end_pm[m] <= lo_col) l = m + 1; else h = m; }
        lo = l;
    }
    const int flo = a.fetch_lo[g];
    auto in_scope = [&](int64_t r) { return r >= lo && r < hi && stepper_pass(rd.flag[r]) && a.end[r] > flo && a.end[r] > rd.pos[r]; };
    auto tag_of = [&](int64_t r) { const int t = a.hp ? a.hp[r] : 0; return (t == 1 || t == 2) ? t : 3; };
    const int one[1] = {centre};

    // centre entry of the row that alignment r OWNS (merged over the later alignments of the same read); 0 = no row
    auto centre_of = [&](int64_t r, int& hp_out) {
        int v = 0; hp_out = 0;
        for (int64_t m = r; m >= 0 && m < hi; m = a.dup_next ? a.dup_next[m] : -1) {
            if (m != r && !in_scope(m)) continue;
            walk_read(rd, a.ck, nullptr, tag_of(m), m, one, 1, [&](int, const Cell& c) { v = c.v; hp_out = c.hp; });
        }
        return v;
    };
    auto owner = [&](int64_t r) {                                  // no earlier in-scope alignment of the same read
        if (!a.dup_prev) return true;
        for (int64_t m = a.dup_prev[r]; m >= lo; m = a.dup_prev[m]) if (in_scope(m)) return false;
        return true;
    };

    // pass 1: rows per HP class
    int cls[3] = {0, 0, 0};
    for (int64_t r0 = lo; r0 < hi; r0 += 32) {
        const int64_t r = r0 + lane;
        int hpv = 0;
        const bool row = in_scope(r) && owner(r) && centre_of(r, hpv) != 0;
#pragma unroll
        for (int t = 0; t < 3; ++t) cls[t] += __popc(__ballot_sync(0xffffffffu, row && hpv == t + 1));
    }
    const int depth = cls[0] + cls[1] + cls[2];
    int base[3] = {0, cls[0], cls[0] + cls[1]};
    int run[3] = {0, 0, 0};
    bool bad = false;
    const bool write = a.hap[0] != nullptr;

    // pass 2: every in-scope alignment counts on its columns; owners of a kept row write it
    for (int64_t r0 = lo; r0 < hi; r0 += 32) {
        const int64_t r = r0 + lane;
        const bool scope = in_scope(r);
        if (scope) {
            const int p0 = rd.pos[r], p1 = a.end[r];
            for (int i = 0; i < nt; ++i) if (tg[i] >= p0 && tg[i] < p1) atomicAdd(&ncol[i], 1);
        }
        const bool own = scope && owner(r);
        int hpv = 0;
        const bool row = own && centre_of(r, hpv) != 0;
        int my_row = -1;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const uint32_t m = __ballot_sync(0xffffffffu, row && hpv == t + 1);
            if (row && hpv == t + 1) my_row = base[t] + run[t] + __popc(m & ((1u << lane) - 1u));
            run[t] += __popc(m);
        }
        if (!own || !write) continue;
        const bool put = write && my_row >= 0 && my_row < a.cap;
        if (put) {
            for (int k = 0; k < 4; ++k) {
                int32_t* h = a.hap[k] + (g * a.cap + my_row) * a.n_hap;
                for (int i = 0; i < a.n_hap; ++i) h[i] = 0;
                int32_t* p = a.pile[k] + (g * a.cap + my_row) * W;
                for (int i = 0; i < W; ++i) p[i] = 0;
            }
        }
        for (int64_t m = r; m >= 0 && m < hi; m = a.dup_next ? a.dup_next[m] : -1) {
            if (m != r && !in_scope(m)) continue;
            bad |= walk_read(rd, a.ck, rd.qual, tag_of(m), m, tg, nt, [&](int i, const Cell& c) {
                if (!put) return;
                const int vals[4] = {c.v, c.hp, c.bq, c.mq};
                for (int k = 0; k < 4; ++k) {
                    if (k == 2 && c.v < 0) continue;              // a deletion leaves the base quality alone (:126-131): an earlier
                                                                  // alignment of the same read may have put one there
                    if (hc[i] >= 0) a.hap[k][(g * a.cap + my_row) * a.n_hap + hc[i]] = vals[k];
                    if (wc[i] >= 0) a.pile[k][(g * a.cap + my_row) * W + wc[i]] = vals[k];
                }
            });
        }
    }
    __syncwarp();
    int fl = __any_sync(0xffffffffu, bad) ? HG_BAD_BASE : 0;
    if (write && depth > a.cap) fl |= HG_ROW_OVERFLOW;
    if (lane == 0) { a.depth[g] = depth; a.gflags[g] = fl; }
    for (int i = lane; i < nt; i += 32) {
        if (hc[i] >= 0) a.n_cols[g * (a.n_hap + W) + hc[i]] = ncol[i];
        if (wc[i] >= 0) a.n_cols[g * (a.n_hap + W) + a.n_hap + wc[i]] = ncol[i];
    }
    if (write) {                                                  // padding rows (write_to_bins.py:14-30: constant -2)
        const int rows = depth < a.cap ? depth : a.cap;
        for (int k = 0; k < 4; ++k) {
            for (int64_t i = (int64_t)rows * a.n_hap + lane; i < (int64_t)a.cap * a.n_hap; i += 32) a.hap[k][g * a.cap * a.n_hap + i] = -2;
            for (int64_t i = (int64_t)rows * W + lane; i < (int64_t)a.cap * W; i += 32) a.pile[k][g * a.cap * W + i] = -2;
        }
    }
}

}  // namespace
}  // namespace nsnp

using namespace nsnp;

extern "C" int64_t nsnp_hap_checkpoint_count(int64_t n_reads, int64_t n_cigar) { return (n_cigar >> kCkShift) + n_reads + 1; }

extern "C" int nsnp_hap_read_ends(const nsnp_reads_t* reads_dev, int32_t* end_dev, int32_t* checkpoints_dev, void* stream)
{
    if (!reads_dev || !end_dev) return set_error(NSNP_E_INVALID, "nsnp_hap_read_ends: null argument");
    if (reads_dev->n_reads == 0) return NSNP_OK;
    read_end_kernel<<<(unsigned)((reads_dev->n_reads + 7) / 8), 256, 0, (cudaStream_t)stream>>>(*reads_dev, end_dev, reinterpret_cast<int2*>(checkpoints_dev));
    return cuda_status("read_end_kernel");
}

extern "C" int nsnp_hap_group_matrices(const nsnp_reads_t* reads_dev, const uint8_t* hp_dev, const int32_t* end_dev, const int32_t* end_pm_dev,
                                       const int32_t* checkpoints_dev,
                                       const int32_t* dup_prev_dev, const int32_t* dup_next_dev, const int32_t* gpos_dev,
                                       const int32_t* fetch_lo_dev, int64_t n_groups, int32_t n_hap, int32_t flank, int32_t cap,
                                       int32_t* n_cols_dev, int32_t* depth_dev, int32_t* flags_dev, int32_t* const* hap_dev,
                                       int32_t* const* pile_dev, void* stream)
{
    if (!reads_dev || !end_dev || !end_pm_dev || !gpos_dev || !fetch_lo_dev || !n_cols_dev || !depth_dev || !flags_dev)
        return set_error(NSNP_E_INVALID, "nsnp_hap_group_matrices: null argument");
    if (n_hap < 1 || n_hap > 31 || !(n_hap & 1) || flank > 32 || n_hap + (flank >= 0 ? 2 * flank + 1 : 0) >= kMaxCols)
        return set_error(NSNP_E_INVALID, "nsnp_hap_group_matrices: n_hap must be odd and < 32, flank <= 32");
    if ((dup_prev_dev == nullptr) != (dup_next_dev == nullptr)) return set_error(NSNP_E_INVALID, "nsnp_hap_group_matrices: dup_prev and dup_next go together");
    if (n_groups <= 0) return NSNP_OK;
    GroupArgs a;
    a.rd = *reads_dev; a.hp = hp_dev; a.end = end_dev; a.end_pm = end_pm_dev; a.ck = reinterpret_cast<const int2*>(checkpoints_dev); a.dup_prev = dup_prev_dev; a.dup_next = dup_next_dev;
    a.gpos = gpos_dev; a.fetch_lo = fetch_lo_dev; a.n_groups = n_groups; a.n_hap = n_hap; a.flank = flank; a.cap = cap;
    a.n_cols = n_cols_dev; a.depth = depth_dev; a.gflags = flags_dev;
    for (int k = 0; k < 4; ++k) { a.hap[k] = hap_dev ? hap_dev[k] : nullptr; a.pile[k] = pile_dev ? pile_dev[k] : nullptr; }
    if (a.hap[0] && (flank < 0 || cap < 1 || !a.pile[0])) return set_error(NSNP_E_INVALID, "nsnp_hap_group_matrices: matrices need flank >= 0, cap >= 1 and both outputs");
    hap_group_kernel<<<(unsigned)((n_groups + kWarps - 1) / kWarps), kWarps * 32, 0, (cudaStream_t)stream>>>(a);
    return cuda_status("hap_group_kernel");
}
