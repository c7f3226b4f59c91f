// s2, VCF record TEXT on the GPU: compact site records (record.cu) -> the bytes predict.py:66-194 writes.
//
// The host text assembly (vcf.cu, ~50 ns of branchy scalar code per record and thread) bounded the end-to-end rate at
// 8 GPUs on 16-32 host cores.  Here one thread formats one record into shared memory, a block scan packs the block's
// records, and the block copies its contiguous text range to global memory; record lengths -> exclusive scan -> write
// is two passes over the records (format twice, store once).  What leaves the device is the text itself (~66 B/site).
//
// Byte identity with the host formatter: every field is integer formatting of values the record already carries.  The
// only libc dependency is the 2-decimal rounding of QUAL when the log-odds sits within 1e-6 of a tie (record.cu flags
// those, ~4 per million): they are listed with their text offset, and nsnp_vcf_text_patch_ties re-evaluates them on the
// host with libc and splices the (rarely different) digits in.
// The `gt_output[ti]` quirk (predict.py:106,119: the BATCH's argmax array indexed by a class index) needs the genotype
// argmax of the first ten sites of the record's 1000-site batch: taken from the record buffer itself when it starts on
// a batch boundary (single-GPU streaming: RegionRunner carries the trailing partial batch into the next region), or from
// a per-contig table (multi-GPU: every rank fills the batches it owns, one small MIN all-reduce completes the table).
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace nsnp {
namespace {

constexpr int kSlot = 152;                 // bytes per record slot: contig (<= 63) + 1 + 88 of fields
constexpr int kTextThreads = 128;           // 128 slots + the packed copy stay under the 48 KB static shared memory limit
constexpr int kMaxContigName = 63;

struct ContigName { char s[kMaxContigName + 1]; int len; };
struct TieEntry { int64_t off; int32_t j; int32_t len; nsnp_site_record_t rec; uint8_t head[10]; uint8_t pad[6]; };   // 64 bytes
static_assert(sizeof(TieEntry) == 64, "tie entry layout");
// deferred batch heads (multi-GPU streaming): the ALT of a fix-up record is one character that depends on the first ten
// sites of its batch, which may live on another rank; the record is written with a placeholder and listed here
struct FixEntry { int64_t off; int32_t j; uint8_t zyo, sref, pad[2]; };                                                  // 16 bytes
static_assert(sizeof(FixEntry) == 16, "fix entry layout");

struct HeadSrc {                            // gt argmax of site ti of the batch that holds global site index g
    const nsnp_site_record_t* rec; int64_t n; int64_t first_index; int64_t batch; const uint8_t* table;
    __host__ __device__ int get(int64_t g, int ti) const {
        const int64_t b = g / batch;
        if (table) return table[b * 10 + ti];
        const int64_t k = b * batch + ti - first_index;
        return (k >= 0 && k < n) ? rec[k].gt : 255;
    }
};

__host__ __device__ inline char* put_dec(char* p, unsigned long long v) {
    char t[20]; int n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = t[--n];
    return p;
}
__host__ __device__ inline int base_idx_hd(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// ALT of a hom / het "fix-up" record (predict.py:101-131): the class whose batch-level argmax is largest among the candidate
// class indices; 0 when the batch has no such site (IndexError -> the record is dropped)
__host__ __device__ inline char fixup_alt(int zyo, char sref, const uint8_t* head)
{
    const char kGt[10][3] = {"AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT"};
    const int tis_hom[4] = {0, 4, 7, 9};
    const int tis_het[6] = {1, 2, 3, 5, 6, 8};
    const int nt = zyo == 1 ? 4 : 6;
    int max_ti = -1, max_v = -1;
    for (int q = 0; q < nt; ++q) {
        const int ti = zyo == 1 ? tis_hom[q] : tis_het[q];
        if (zyo == 1 && kGt[ti][0] == sref) continue;
        const int v = head[ti];
        if (v == 255) return 0;                                               // IndexError on the batch argmax array
        if (v > max_v) { max_v = v; max_ti = ti; }
    }
    return zyo == 1 ? kGt[max_ti][0] : (kGt[max_ti][0] == sref ? kGt[max_ti][1] : kGt[max_ti][0]);
}

// One record (predict.py:66-194 after the numeric half): returns the text length, 0 when predict.py writes nothing.
// head[ti] = gt argmax of site ti of the record's batch, 255 when the batch has no such site (IndexError -> dropped).
// head == nullptr: deferred -- a fix-up record gets the placeholder ALT '?' and *alt_off = its offset in the record.
__host__ __device__ inline int format_record(char* out, const char* contig, int clen, const nsnp_site_record_t& r, const uint8_t* head,
                                             long long gt_q, long long zy_q, int* alt_off = nullptr)
{
    const char kGt[10][3] = {"AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT"};       // options.py:8-17
    bool placeholder = false;
    const char kZy[3][4] = {"0/0", "1/1", "0/1"};                                              // options.py:30
    const int gt = r.gt, zyo = r.zy;
    if (gt >= 10) return 0;                                                   // predict.py:68
    if (r.flags & NSNP_REC_DROP) return 0;                                    // calculate_score raised
    const char sref = (char)r.ref;
    char alt[4]; int na = 0;
    for (int k = 0; k < 2; ++k) if (kGt[gt][k] != sref) alt[na++] = kGt[gt][k];
    const long long qual = gt_q < zy_q ? gt_q : zy_q;
    char alts[4] = {0, 0, 0, 0}; int nalt = 0;
    const char* zy = kZy[zyo];
    long long q100; bool refcall = false;
    if (na == 0) {
        if (zyo == 0) { alts[0] = sref; nalt = 1; q100 = qual; refcall = true; }
        else {
            if (head) { alts[0] = fixup_alt(zyo, sref, head); if (!alts[0]) return 0; }
            else { alts[0] = '?'; placeholder = true; }
            nalt = 1; q100 = zy_q;
        }
    } else {
        if (na == 1 || alt[0] == alt[1]) { alts[0] = alt[0]; nalt = 1; }
        else { alts[0] = alt[0]; alts[1] = ','; alts[2] = alt[1]; nalt = 3; }
        if (nalt == 3 && zyo != 2) zy = "1/2";
        q100 = zyo == 0 ? gt_q : qual;
    }
    char* p = out;
    for (int i = 0; i < clen; ++i) *p++ = contig[i];
    *p++ = '\t';
    p = put_dec(p, (unsigned long long)(uint32_t)r.pos1);
    *p++ = '\t'; *p++ = '.'; *p++ = '\t'; *p++ = sref; *p++ = '\t';
    if (alt_off) *alt_off = placeholder ? (int)(p - out) : -1;
    for (int i = 0; i < nalt; ++i) *p++ = alts[i];
    *p++ = '\t';
    const unsigned long long ip = (unsigned long long)(q100 / 100); const int f2 = (int)(q100 - (long long)ip * 100);
    p = put_dec(p, ip); *p++ = '.'; *p++ = (char)('0' + f2 / 10); if (f2 % 10) *p++ = (char)('0' + f2 % 10);
    *p++ = '\t';
    if (refcall) { const char f[] = "RefCall"; for (int i = 0; i < 7; ++i) *p++ = f[i]; }
    else { const char f[] = "PASS"; for (int i = 0; i < 4; ++i) *p++ = f[i]; }
    { const char f[] = "\t.\tGT:GQ:DP:AF\t"; for (int i = 0; i < 15; ++i) *p++ = f[i]; }
    *p++ = zy[0]; *p++ = zy[1]; *p++ = zy[2]; *p++ = ':';
    p = put_dec(p, ip);
    *p++ = ':';
    if (r.depth < 0) { *p++ = '-'; p = put_dec(p, (unsigned long long)(-(long long)r.depth)); } else p = put_dec(p, (unsigned long long)r.depth);
    *p++ = ':';
    if (r.af_q == NSNP_AF_ONE) { const char f[] = "1.000000"; for (int i = 0; i < 8; ++i) *p++ = f[i]; }
    else if (r.af_q == NSNP_AF_NAN) { *p++ = 'n'; *p++ = 'a'; *p++ = 'n'; }
    else if (r.af_q == NSNP_AF_NEG_INF) { *p++ = '-'; *p++ = 'i'; *p++ = 'n'; *p++ = 'f'; }
    else {
        int32_t aq = r.af_q;
        if (aq <= -3) { *p++ = '-'; aq = -(aq + 3); }
        const uint32_t u = (uint32_t)aq, ipa = u / 1000000u; uint32_t f = u - ipa * 1000000u;
        p = put_dec(p, ipa); *p++ = '.';
        char d[6]; for (int i = 5; i >= 0; --i) { d[i] = (char)('0' + f % 10); f /= 10; }
        for (int i = 0; i < 6; ++i) *p++ = d[i];
    }
    *p++ = '\n';
    return (int)(p - out);
}

// pass 0 (kWrite = false): block text sizes.  pass 1: format again, pack the block's records in shared memory, copy out.
template <bool kWrite>
__global__ void __launch_bounds__(kTextThreads) vcf_text_kernel(const nsnp_site_record_t* __restrict__ rec, int64_t n, const int32_t* __restrict__ n_dev,
                                                               HeadSrc hs, ContigName name, int64_t* __restrict__ block_off,
                                                               char* __restrict__ text, int64_t capacity, TieEntry* __restrict__ ties,
                                                               int32_t* __restrict__ tie_count, int tie_cap, int deferred,
                                                               FixEntry* __restrict__ fixes, int32_t* __restrict__ fix_count)
{
    __shared__ __align__(16) char slot[kTextThreads][kSlot];
    __shared__ int warp_tot[kTextThreads / 32];
    if (n_dev) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
    hs.n = n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t j = (int64_t)blockIdx.x * kTextThreads + tid;
    int len = 0;
    nsnp_site_record_t r;
    uint8_t head[10];
    int alt_off = -1;
    if (j < n) {
        r = rec[j];
        const bool fix = !deferred && r.gt < 10 && !(r.flags & NSNP_REC_DROP) && r.zy != 0;      // only fix-up records read the batch heads
        for (int k = 0; k < 10; ++k) head[k] = fix ? (uint8_t)hs.get(hs.first_index + j, k) : (uint8_t)255;
        len = format_record(slot[tid], name.s, name.len, r, deferred ? nullptr : head, r.q100_gt, r.q100_zy, &alt_off);
    }
    int inc = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int off = inc - len, total = 0;
    for (int w = 0; w < kTextThreads / 32; ++w) { if (w < warp) off += warp_tot[w]; total += warp_tot[w]; }
    if (!kWrite) { if (tid == 0) block_off[blockIdx.x] = total; return; }
    const int64_t base = block_off[blockIdx.x];
    if (base + total > capacity) return;                        // reported by the host wrapper from the total length
    if (len && (r.flags & (NSNP_REC_TIE_GT | NSNP_REC_TIE_ZY))) {
        const int t = atomicAdd(tie_count, 1);
        if (t < tie_cap) {
            TieEntry e; e.off = base + off; e.j = (int32_t)j; e.len = len; e.rec = r;
            for (int k = 0; k < 10; ++k) e.head[k] = head[k];
            for (int k = 0; k < 6; ++k) e.pad[k] = 0;
            ties[t] = e;
        }
    }
    if (len && alt_off >= 0) {
        const int t = atomicAdd(fix_count, 1);
        FixEntry e; e.off = base + off + alt_off; e.j = (int32_t)j; e.zyo = r.zy; e.sref = r.ref; e.pad[0] = e.pad[1] = 0;
        fixes[t] = e;                                           // capacity = n: every record could be a fix-up
    }
    // pack: this thread's record moves to its packed position in a second buffer (reuse of the slots is not possible in
    // place: packed ranges overlap other threads' unread slots), then the block stores its range with coalesced bytes
    __shared__ __align__(16) char packed[kTextThreads * kSlot / 2];      // a block's records average ~70 bytes: half the slots
    const bool fits = total <= (int)sizeof(packed);
    if (fits) {
        for (int i = 0; i < len; ++i) packed[off + i] = slot[tid][i];
        __syncthreads();
        for (int i = tid; i < total; i += kTextThreads) text[base + i] = packed[i];
    } else {
        for (int i = 0; i < len; ++i) text[base + off + i] = slot[tid][i];
    }
}

__global__ void __launch_bounds__(1024) block_scan_kernel(int64_t* __restrict__ v, int64_t nb, int64_t* __restrict__ total_out)
{
    __shared__ int64_t part[1024];
    const int tid = threadIdx.x;
    const int64_t per = (nb + 1023) / 1024;
    const int64_t lo = min(nb, tid * per), hi = min(nb, lo + per);
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += v[i];
    part[tid] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int64_t t = tid >= d ? part[tid - d] : 0;
        __syncthreads();
        part[tid] += t;
        __syncthreads();
    }
    int64_t run = tid ? part[tid - 1] : 0;
    for (int64_t i = lo; i < hi; ++i) { const int64_t x = v[i]; v[i] = run; run += x; }
    if (tid == 1023) { v[nb] = part[1023]; if (total_out) *total_out = part[1023]; }
}

__global__ void __launch_bounds__(256) batch_heads_kernel(const nsnp_site_record_t* __restrict__ rec, int64_t n, const int32_t* __restrict__ n_dev,
                                                          int64_t first_index, int64_t batch, uint8_t* __restrict__ heads)
{
    if (n_dev) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
    // only the first ten sites of every batch matter: one thread per (batch, k)
    const int64_t b0 = first_index / batch, b1 = (first_index + n + batch - 1) / batch;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = b0 + t / 10; const int k = (int)(t % 10);
    if (b >= b1) return;
    const int64_t j = b * batch + k - first_index;
    if (j >= 0 && j < n) heads[b * 10 + k] = rec[j].gt;
}

struct TextWs { int64_t* block_off; int32_t* tie_count; TieEntry* ties; int32_t* fix_count; FixEntry* fixes; };
constexpr int kTieCap = 4096;
inline size_t carve_text(void* base, int64_t n, TextWs* w) {
    const int64_t nb = (n + kTextThreads - 1) / kTextThreads;
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base ? (char*)base + off : nullptr; off += (bytes + 255) / 256 * 256; return p; };
    w->block_off = (int64_t*)take((size_t)(nb + 2) * 8);
    w->tie_count = (int32_t*)take(64);
    w->ties = (TieEntry*)take((size_t)kTieCap * sizeof(TieEntry));
    w->fix_count = (int32_t*)take(64);
    w->fixes = (FixEntry*)take((size_t)(n + 1) * sizeof(FixEntry));
    return off;
}

// calculate_score (predict.py:31-34) with libc, as vcf.cu evaluates it; q100 = round(x, 2) * 100
bool host_q100(float p, long long* q100) {
    const float a = 1.0f - p;
    const float r = a / p;
    const double x = (double)r;
    if (!(x > 0.0)) return false;
    static const double kScale = -10.0 * (1.0 / log(10.0));
    volatile double t = kScale * log(x);
    t = t + 10.0;
    const double v = t > 0.0 ? t : 0.0;
    char buf[64];
    snprintf(buf, sizeof buf, "%.2f", v);
    *q100 = (long long)floor(strtod(buf, nullptr) * 100.0 + 0.5);
    return true;
}

}  // namespace
}  // namespace nsnp

using namespace nsnp;

extern "C" {

size_t nsnp_vcf_text_workspace_bytes(int64_t n) { TextWs w; return carve_text(nullptr, n < 0 ? 0 : n, &w); }
int64_t nsnp_vcf_text_capacity(int64_t n, const char* contig) { return n * (int64_t)(88 + (contig ? strlen(contig) : kMaxContigName)) + 256; }

static int text_records_impl(const char* contig, const nsnp_site_record_t* rec_dev, int64_t n, const int32_t* n_dev, int64_t first_index,
                             int64_t batch_size, const uint8_t* heads_dev, char* text_dev, int64_t text_capacity, int64_t* text_len_dev,
                             void* workspace_dev, size_t workspace_bytes, int deferred, void* stream_);

int nsnp_vcf_text_records(const char* contig, const nsnp_site_record_t* rec_dev, int64_t n, const int32_t* n_dev, int64_t first_index,
                          int64_t batch_size, const uint8_t* heads_dev, char* text_dev, int64_t text_capacity, int64_t* text_len_dev,
                          void* workspace_dev, size_t workspace_bytes, void* stream_)
{
    return text_records_impl(contig, rec_dev, n, n_dev, first_index, batch_size, heads_dev, text_dev, text_capacity, text_len_dev, workspace_dev,
                             workspace_bytes, 0, stream_);
}

int nsnp_vcf_text_records_deferred(const char* contig, const nsnp_site_record_t* rec_dev, int64_t n, const int32_t* n_dev, char* text_dev,
                                   int64_t text_capacity, int64_t* text_len_dev, void* workspace_dev, size_t workspace_bytes, void* stream_)
{
    return text_records_impl(contig, rec_dev, n, n_dev, 0, 1000, nullptr, text_dev, text_capacity, text_len_dev, workspace_dev, workspace_bytes, 1, stream_);
}

static int text_records_impl(const char* contig, const nsnp_site_record_t* rec_dev, int64_t n, const int32_t* n_dev, int64_t first_index,
                             int64_t batch_size, const uint8_t* heads_dev, char* text_dev, int64_t text_capacity, int64_t* text_len_dev,
                             void* workspace_dev, size_t workspace_bytes, int deferred, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!contig || n < 0 || batch_size <= 0 || first_index < 0 || !text_len_dev || (n > 0 && (!rec_dev || !text_dev || !workspace_dev)))
        return set_error(NSNP_E_INVALID, "nsnp_vcf_text_records: bad argument");
    if (!deferred && !heads_dev && first_index % batch_size != 0)
        return set_error(NSNP_E_INVALID, "nsnp_vcf_text_records: without a batch-head table the records must start on a batch boundary");
    ContigName name;
    const size_t clen = strlen(contig);
    if (clen > kMaxContigName) return set_error(NSNP_E_UNSUPPORTED, "contig name longer than %d characters", kMaxContigName);
    memset(&name, 0, sizeof name); memcpy(name.s, contig, clen); name.len = (int)clen;
    if (nsnp_device_count() == 0) return set_error(NSNP_E_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (n == 0) { cudaMemsetAsync(text_len_dev, 0, 8, stream); return cuda_status("nsnp_vcf_text_records"); }
    TextWs w;
    if (carve_text(workspace_dev, n, &w) > workspace_bytes) return set_error(NSNP_E_WORKSPACE, "vcf text workspace too small");
    const int64_t nb = (n + kTextThreads - 1) / kTextThreads;
    HeadSrc hs{rec_dev, n, first_index, batch_size, heads_dev};
    cudaMemsetAsync(w.tie_count, 0, 4, stream);
    cudaMemsetAsync(w.fix_count, 0, 4, stream);
    ProfScope prof(NSNP_PROF_VCF_TEXT, stream);
    vcf_text_kernel<false><<<(unsigned)nb, kTextThreads, 0, stream>>>(rec_dev, n, n_dev, hs, name, w.block_off, nullptr, 0, nullptr, nullptr, 0, deferred, nullptr, nullptr);
    block_scan_kernel<<<1, 1024, 0, stream>>>(w.block_off, nb, text_len_dev);
    vcf_text_kernel<true><<<(unsigned)nb, kTextThreads, 0, stream>>>(rec_dev, n, n_dev, hs, name, w.block_off, text_dev, text_capacity, w.ties, w.tie_count, kTieCap, deferred, w.fixes, w.fix_count);
    return cuda_status("vcf_text_kernel");
}

/* the tie list of the last nsnp_vcf_text_records call on this workspace: count (int32) followed, 64 bytes in, by 64-byte entries */
int nsnp_vcf_text_ties(const void* workspace_dev, int64_t n, const void** count_dev, const void** entries_dev, int32_t* capacity)
{
    TextWs w; carve_text(const_cast<void*>(workspace_dev), n < 0 ? 0 : n, &w);
    if (count_dev) *count_dev = w.tie_count;
    if (entries_dev) *entries_dev = w.ties;
    if (capacity) *capacity = kTieCap;
    return NSNP_OK;
}

/* the fix-up list of the last nsnp_vcf_text_records_deferred call on this workspace: int32 count + 16-byte entries */
int nsnp_vcf_text_fixups(const void* workspace_dev, int64_t n, const void** count_dev, const void** entries_dev)
{
    TextWs w; carve_text(const_cast<void*>(workspace_dev), n < 0 ? 0 : n, &w);
    if (count_dev) *count_dev = w.fix_count;
    if (entries_dev) *entries_dev = w.fixes;
    return NSNP_OK;
}

/* host: writes the real ALT character of every listed fix-up record once the batch-head table of its contig is complete.
 * first_index: contig-wide index of the region's first record.  *n_drop counts records whose batch has no such site
 * (predict.py drops them; only possible in the last, short batch of a contig): the caller re-formats that region exactly. */
int nsnp_vcf_text_patch_heads(char* text_host, int64_t text_len, const void* fix_host, int32_t n_fix, int64_t first_index, int64_t batch_size,
                              const uint8_t* heads_table, int32_t* n_drop)
{
    if (!text_host || n_fix < 0 || (n_fix > 0 && (!fix_host || !heads_table)) || batch_size <= 0) return set_error(NSNP_E_INVALID, "nsnp_vcf_text_patch_heads: bad argument");
    const FixEntry* f = (const FixEntry*)fix_host;
    int32_t drops = 0;
    for (int32_t i = 0; i < n_fix; ++i) {
        if (f[i].off < 0 || f[i].off >= text_len) return set_error(NSNP_E_INVALID, "nsnp_vcf_text_patch_heads: offset out of range");
        const char a = fixup_alt(f[i].zyo, (char)f[i].sref, heads_table + ((first_index + f[i].j) / batch_size) * 10);
        if (!a) ++drops; else text_host[f[i].off] = a;
    }
    if (n_drop) *n_drop = drops;
    return NSNP_OK;
}

int nsnp_vcf_batch_heads(const nsnp_site_record_t* rec_dev, int64_t n, const int32_t* n_dev, int64_t first_index, int64_t batch_size,
                         uint8_t* heads_dev, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || batch_size < 10 || first_index < 0 || (n > 0 && (!rec_dev || !heads_dev))) return set_error(NSNP_E_INVALID, "nsnp_vcf_batch_heads: bad argument");
    if (n == 0) return NSNP_OK;
    if (nsnp_device_count() == 0) return set_error(NSNP_E_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    const int64_t nbatch = (first_index + n + batch_size - 1) / batch_size - first_index / batch_size;
    batch_heads_kernel<<<(unsigned)((nbatch * 10 + 255) / 256), 256, 0, stream>>>(rec_dev, n, n_dev, first_index, batch_size, heads_dev);
    return cuda_status("batch_heads_kernel");
}

/* host: re-evaluates the flagged records with libc and splices differing bytes into the text; returns the new length
 * (or the negative of the capacity needed).  ties_host: n_ties 64-byte entries copied from the device list. */
int64_t nsnp_vcf_text_patch_ties_at(const char* contig, char* text_host, int64_t text_len, int64_t text_capacity, const void* ties_host, int32_t n_ties,
                                    int64_t first_index, int64_t batch_size, const uint8_t* heads_table);

int64_t nsnp_vcf_text_patch_ties(const char* contig, char* text_host, int64_t text_len, int64_t text_capacity, const void* ties_host, int32_t n_ties)
{
    return nsnp_vcf_text_patch_ties_at(contig, text_host, text_len, text_capacity, ties_host, n_ties, 0, 1, nullptr);
}

/* heads_table != NULL (deferred text): the listed records carry no batch heads; they are looked up in the contig's table */
int64_t nsnp_vcf_text_patch_ties_at(const char* contig, char* text_host, int64_t text_len, int64_t text_capacity, const void* ties_host, int32_t n_ties,
                                    int64_t first_index, int64_t batch_size, const uint8_t* heads_table)
{
    if (!contig || !text_host || text_len < 0 || n_ties < 0 || (n_ties > 0 && !ties_host) || batch_size <= 0) return 0;
    const int clen = (int)strlen(contig);
    std::vector<TieEntry> t((const TieEntry*)ties_host, (const TieEntry*)ties_host + n_ties);
    std::sort(t.begin(), t.end(), [](const TieEntry& a, const TieEntry& b) { return a.off > b.off; });     // back to front: offsets stay valid
    for (const TieEntry& e : t) {
        long long gq = e.rec.q100_gt, zq = e.rec.q100_zy;
        bool ok = true;
        if (e.rec.flags & NSNP_REC_TIE_GT) ok = ok && host_q100(e.rec.p_gt, &gq);
        if (e.rec.flags & NSNP_REC_TIE_ZY) ok = ok && host_q100(e.rec.p_zy, &zq);
        char buf[kSlot + 64];
        const uint8_t* head = heads_table ? heads_table + ((first_index + e.j) / batch_size) * 10 : e.head;
        const int len = ok ? format_record(buf, contig, clen, e.rec, head, gq, zq) : 0;
        if (e.off < 0 || e.off + e.len > text_len) return 0;
        if (len == e.len && memcmp(buf, text_host + e.off, (size_t)len) == 0) continue;
        const int64_t new_len = text_len + (len - e.len);
        if (new_len > text_capacity) return -new_len;
        memmove(text_host + e.off + len, text_host + e.off + e.len, (size_t)(text_len - e.off - e.len));
        memcpy(text_host + e.off, buf, (size_t)len);
        text_len = new_len;
    }
    return text_len;
}

/* host reference of the device formatter (same code path compiled for the host): records [first_index, first_index + n) */
int64_t nsnp_vcf_format_records_at(const char* contig, const nsnp_site_record_t* rec, int64_t n, int64_t first_index, int64_t batch_size,
                                   const uint8_t* heads, char* out, int64_t out_capacity)
{
    if (!contig || n < 0 || batch_size <= 0 || (n > 0 && (!rec || !out))) return 0;
    if (!heads && first_index % batch_size != 0) return 0;
    const int clen = (int)strlen(contig);
    HeadSrc hs{rec, n, first_index, batch_size, heads};
    int64_t o = 0;
    char buf[kSlot + 256];
    for (int64_t j = 0; j < n; ++j) {
        uint8_t head[10];
        for (int k = 0; k < 10; ++k) head[k] = (uint8_t)hs.get(first_index + j, k);
        long long gq = rec[j].q100_gt, zq = rec[j].q100_zy;
        bool ok = true;
        if (rec[j].gt < 10 && !(rec[j].flags & NSNP_REC_DROP)) {
            if (rec[j].flags & NSNP_REC_TIE_GT) ok = ok && host_q100(rec[j].p_gt, &gq);
            if (rec[j].flags & NSNP_REC_TIE_ZY) ok = ok && host_q100(rec[j].p_zy, &zq);
        }
        const int len = ok ? format_record(buf, contig, clen, rec[j], head, gq, zq) : 0;
        if (o + len <= out_capacity) memcpy(out + o, buf, (size_t)len);
        o += len;
    }
    return o <= out_capacity ? o : -o;
}

}  // extern "C"
