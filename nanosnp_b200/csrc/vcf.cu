// s2 host side: VCF record text for one batch of sites -- a native restatement of the per-site loop of
// PileupModel/predict.py:54-194, including the behaviours that define "identical VCF" (SURVEY 8a, P13):
//   * `gt_output[ti]` indexes the BATCH argmax array with a class index (predict.py:106,119,150,163), so the
//     ALT of hom/het fix-up records depends on the first ten sites of the batch, and a batch with <= ti
//     sites raises IndexError -> the record is silently dropped by the bare `except` (predict.py:193);
//   * calculate_score (predict.py:31-34) runs in float32 under NumPy >= 2, so p == 1.0 gives log(0) ->
//     ValueError -> record dropped;
//   * QUAL is str(round(x, 2)) of a Python float, AF is '%f' of a float32 quotient, DP is '%d' of a float32.
// Host code only (no kernels): this is the caller either side of the GPU path.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace {

const char* const kGt[21] = {"AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT", "DD",
                             "AD", "CD", "GD", "TD", "II", "AI", "CI", "GI", "TI", "ID"};     // options.py:8-28
const char* const kZy[3] = {"0/0", "1/1", "0/1"};                                              // options.py:30

inline int base_idx(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// calculate_score, predict.py:31-34.  Returns false when Python would raise (log of 0).
bool calc_score(float p, double* out) {
    const float a = 1.0f - p;            // (1.0 - p) + 1e-300 stays float32 under NumPy 2; 1e-300 -> 0.0f
    const float r = a / p;
    const double x = (double)r;
    if (!(x > 0.0)) return false;        // math.log(0.0) / log(negative) / log(nan): ValueError
    static const double kScale = -10.0 * (1.0 / log(10.0));     // -10 * log(e, 10) = -10 * (log(e) / log(10)), log(e) == 1.0
    volatile double t = kScale * log(x);
    t = t + 10.0;
    double v = t > 0.0 ? t : 0.0;        // max(tmp, 0)
    char buf[64];
    snprintf(buf, sizeof buf, "%.2f", v);          // round(tmp, 2): correctly rounded decimal, half-even on exact ties
    *out = strtod(buf, nullptr);
    return true;
}

// str(float) for a value that is the nearest double of a 2-decimal number
int fmt_pyfloat(char* dst, double v) {
    char buf[64];
    int n = snprintf(buf, sizeof buf, "%.2f", v);
    while (n > 0 && buf[n - 1] == '0' && buf[n - 2] != '.') --n;
    memcpy(dst, buf, (size_t)n);
    return n;
}

struct Out {
    char* p; int64_t cap; int64_t n;
    void put(const char* s, size_t len) { if (n + (int64_t)len <= cap) memcpy(p + n, s, len); n += (int64_t)len; }
    void puts_(const char* s) { put(s, strlen(s)); }
};

void write_record(Out& o, const char* contig, long long pos, char ref, const char* alt, double q, const char* filter,
                  const char* zy, float depth, float af, bool af_is_one)
{
    char line[512], qs[64];
    const int qn = fmt_pyfloat(qs, q); qs[qn] = 0;
    char afs[64];
    const double afd = af_is_one ? 1.0 : (double)af;
    if (isnan(afd)) strcpy(afs, "nan"); else if (isinf(afd)) strcpy(afs, afd > 0 ? "inf" : "-inf"); else snprintf(afs, sizeof afs, "%f", afd);
    const int n = snprintf(line, sizeof line, "%s\t%lld\t.\t%c\t%s\t%s\t%s\t.\tGT:GQ:DP:AF\t%s:%lld:%lld:%s\n",
                           contig, pos, ref, alt, qs, filter, zy, (long long)q, (long long)depth, afs);
    o.put(line, (size_t)n);
}

}  // namespace

extern "C" int64_t nsnp_vcf_format_batch(const char* contig, int64_t n, const int32_t* pos1, const uint8_t* refbase,
                                         const float* gt_prob, const float* zy_prob, const float* cov8,
                                         char* out, int64_t out_capacity)
{
    if (!contig || n < 0 || (n > 0 && (!pos1 || !refbase || !gt_prob || !zy_prob || !cov8))) return 0;
    Out o{out, out ? out_capacity : 0, 0};
    // batch-level argmax arrays (predict.py:56-57)
    int head_gt[10];
    const int nhead = n < 10 ? (int)n : 10;
    auto argmax = [](const float* v, int m) { int b = 0; for (int i = 1; i < m; ++i) if (v[i] > v[b]) b = i; return b; };
    for (int i = 0; i < nhead; ++i) head_gt[i] = argmax(gt_prob + (size_t)i * 21, 21);

    for (int64_t j = 0; j < n; ++j) {
        const float* gp = gt_prob + j * 21; const float* zp = zy_prob + j * 3;
        const int gt = argmax(gp, 21), zyo = argmax(zp, 3);
        if (gt >= 10) continue;                                               // predict.py:68
        const char sref = (char)refbase[j];
        const char* label = kGt[gt];
        const char* zy = kZy[zyo];
        const float* cov = cov8 + j * 8;
        float neg = 0.f; for (int k = 0; k < 8; ++k) if (cov[k] < 0.f) neg += cov[k];
        const float depth = -1.0f * neg;                                      // predict.py:76
        // alt = label minus every occurrence of sref (predict.py:78,90)
        char alt[4]; int na = 0;
        for (int k = 0; k < 2; ++k) if (label[k] != sref) alt[na++] = label[k];
        alt[na] = 0;
        float support = 0.f;
        for (int k = 0; k < na; ++k) { const int b = base_idx(alt[k]); if (b < 0) goto next_site; support += cov[b]; support += cov[b + 4]; }
        {
            float af = support / depth;                                       // float32 quotient (0/0 -> nan, x/0 -> inf)
            bool af_one = false;
            if (af > 1.0f) af_one = true;                                     // predict.py:83-84
            double gt_q, zy_q;
            if (!calc_score(gp[gt], &gt_q)) continue;                         // ValueError -> except: continue
            if (!calc_score(zp[zyo], &zy_q)) continue;
            const double qual = gt_q < zy_q ? gt_q : zy_q;

            if (na == 0) {
                if (zyo == 0) {
                    const char a1[2] = {sref, 0};
                    write_record(o, contig, pos1[j], sref, a1, qual, "RefCall", zy, depth, af, af_one);
                } else if (zyo == 1) {
                    static const int tis[4] = {0, 4, 7, 9};
                    int max_ti = -1, max_v = -1; bool raised = false;
                    for (int q = 0; q < 4; ++q) {
                        const int ti = tis[q];
                        if (kGt[ti][0] == sref) continue;
                        if (ti >= n) { raised = true; break; }                // IndexError on the batch array
                        if (head_gt[ti] > max_v) { max_v = head_gt[ti]; max_ti = ti; }
                    }
                    if (raised) continue;
                    const char a1[2] = {kGt[max_ti][0], 0};
                    write_record(o, contig, pos1[j], sref, a1, zy_q, "PASS", zy, depth, af, af_one);
                } else {
                    static const int tis[6] = {1, 2, 3, 5, 6, 8};
                    int max_ti = -1, max_v = -1; bool raised = false;
                    for (int q = 0; q < 6; ++q) {
                        const int ti = tis[q];
                        if (ti >= n) { raised = true; break; }
                        if (head_gt[ti] > max_v) { max_v = head_gt[ti]; max_ti = ti; }
                    }
                    if (raised) continue;
                    const char a1[2] = {kGt[max_ti][0] == sref ? kGt[max_ti][1] : kGt[max_ti][0], 0};
                    write_record(o, contig, pos1[j], sref, a1, zy_q, "PASS", zy, depth, af, af_one);
                }
                continue;
            }
            char alts[8];
            if (na == 1) { alts[0] = alt[0]; alts[1] = 0; }
            else if (alt[0] == alt[1]) { alts[0] = alt[0]; alts[1] = 0; }     // "CC" -> "C"
            else { alts[0] = alt[0]; alts[1] = ','; alts[2] = alt[1]; alts[3] = 0; }
            if (strlen(alts) >= 3 && zyo != 2) zy = "1/2";                    // predict.py:140-141
            // predict.py:143-176 (`alt == sref and zy != 0`) cannot fire: alt never contains sref
            if (zyo == 0) {                                                   // predict.py:177-185
                write_record(o, contig, pos1[j], sref, alts, gt_q, "PASS", zy, depth, af, af_one);
                continue;
            }
            write_record(o, contig, pos1[j], sref, alts, qual, "PASS", zy, depth, af, af_one);
        }
    next_site:;
    }
    if (o.n > o.cap) return -o.n;
    return o.n;
}


// ===================================================================================================
// Fast path: same records, hand-rolled number formatting, batches formatted on host threads.
// Every shortcut falls back to the libc path above whenever its exactness argument does not hold.
// ===================================================================================================
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// decimal digits two at a time from a 200-byte table, 32-bit arithmetic (every VCF field here fits 32 bits; larger values
// peel off 9-digit groups first).  This is most of the per-record cost of the text assembly.
alignas(64) const char kDigits2[201] =
    "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
    "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
inline char* put_u32(char* p, uint32_t v) {
    if (v < 100) {
        if (v < 10) { *p++ = (char)('0' + v); return p; }
        memcpy(p, kDigits2 + 2 * v, 2); return p + 2;
    }
    char tmp[20]; int n = 10;                                      // digits end at tmp[10]; the tail keeps the fixed-size copy in bounds
    while (v >= 100) { const uint32_t q = v / 100; n -= 2; memcpy(tmp + n, kDigits2 + 2 * (v - q * 100), 2); v = q; }
    if (v >= 10) { n -= 2; memcpy(tmp + n, kDigits2 + 2 * v, 2); } else tmp[--n] = (char)('0' + v);
    memcpy(p, tmp + n, 10 - n > 8 ? 10 : 8);                      // over-copy a fixed size: the caller's buffer has slack
    return p + (10 - n);
}
inline char* put_uint(char* p, unsigned long long v) {
    if (v <= 0xFFFFFFFFull) return put_u32(p, (uint32_t)v);
    const unsigned long long hi = v / 1000000000ull; const uint32_t lo = (uint32_t)(v - hi * 1000000000ull);
    p = put_uint(p, hi);
    char d[9]; uint32_t f = lo; for (int i = 8; i >= 0; --i) { d[i] = (char)('0' + f % 10); f /= 10; }
    memcpy(p, d, 9); return p + 9;
}

// QUAL: str(float(round(v, 2))) and int(...) of it.  v >= 0.  round() is correct rounding of the exact binary value
// to 2 decimals (half-even on exact ties); v*100 is inexact, so anything within 1e-6 of a tie goes the libc way.
inline char* put_qual(char* p, double v, long long* int_part) {
    const double t = v * 100.0;
    const double fl = floor(t);
    const double fr = t - fl;
    long long q100;
    if (fabs(fr - 0.5) < 1e-6 || !(t < 9.0e15)) {
        char buf[64]; snprintf(buf, sizeof buf, "%.2f", v);
        const double r = strtod(buf, nullptr);
        q100 = (long long)floor(r * 100.0 + 0.5);
    } else {
        q100 = (long long)fl + (fr > 0.5 ? 1 : 0);
    }
    const long long ip = q100 / 100; const int f2 = (int)(q100 % 100);
    *int_part = ip;
    p = put_uint(p, (unsigned long long)ip);
    *p++ = '.';
    *p++ = (char)('0' + f2 / 10);
    if (f2 % 10) *p++ = (char)('0' + f2 % 10);
    return p;
}

// '%f' % af: af is a float32 quotient in [0, 1] (or nan); float32 * 1e6 is exact in double, so the correctly rounded
// 6-decimal value (half-even on exact ties, as printf does) follows from the exact product.
inline char* put_af(char* p, float af, bool af_one) {
    if (af_one) { memcpy(p, "1.000000", 8); return p + 8; }
    if (af != af) { memcpy(p, "nan", 3); return p + 3; }
    const double r = (double)af * 1.0e6;
    if (!(r >= 0.0) || r > 1.0e6) { const int n = snprintf(p, 32, "%f", (double)af); return p + n; }
    double q = floor(r); const double fr = r - q;
    if (fr > 0.5 || (fr == 0.5 && fmod(q, 2.0) == 1.0)) q += 1.0;
    const unsigned long long u = (unsigned long long)q;
    p = put_uint(p, u / 1000000ull);
    *p++ = '.';
    unsigned long long f = u % 1000000ull;
    char d[6]; for (int i = 5; i >= 0; --i) { d[i] = (char)('0' + f % 10); f /= 10; }
    memcpy(p, d, 6);
    return p + 6;
}

inline bool calc_score_fast(float p, double* out) {
    const float a = 1.0f - p;
    const float r = a / p;
    const double x = (double)r;
    if (!(x > 0.0)) return false;
    static const double kScale = -10.0 * (1.0 / log(10.0));
    volatile double t = kScale * log(x);
    t = t + 10.0;
    *out = t > 0.0 ? t : 0.0;          // unrounded: put_qual rounds
    return true;
}

inline char* put_record(char* p, const char* contig, size_t clen, long long pos, char ref, const char* alt, double q, const char* filter,
                        const char* zy, float depth, float af, bool af_one)
{
    memcpy(p, contig, clen); p += clen; *p++ = '\t';
    p = put_uint(p, (unsigned long long)pos);
    *p++ = '\t'; *p++ = '.'; *p++ = '\t'; *p++ = ref; *p++ = '\t';
    for (const char* a = alt; *a; ++a) *p++ = *a;
    *p++ = '\t';
    long long qi = 0;
    p = put_qual(p, q, &qi);
    *p++ = '\t';
    for (const char* a = filter; *a; ++a) *p++ = *a;
    memcpy(p, "\t.\tGT:GQ:DP:AF\t", 15); p += 15;
    for (const char* a = zy; *a; ++a) *p++ = *a;
    *p++ = ':';
    p = put_uint(p, (unsigned long long)qi);
    *p++ = ':';
    const long long dp = (long long)depth;
    if (dp < 0) { *p++ = '-'; p = put_uint(p, (unsigned long long)(-dp)); } else p = put_uint(p, (unsigned long long)dp);
    *p++ = ':';
    p = put_af(p, af, af_one);
    *p++ = '\n';
    return p;
}

// one batch (predict.py:54-194), appended to `o`
// writes the records at p (the caller provides kMaxRecordBytes + clen bytes per site) and returns the end
char* format_batch_fast(char* p, const char* contig, size_t clen, int64_t n, const int32_t* pos1, const uint8_t* refbase,
                       const float* gt_prob, const float* zy_prob, const float* cov8)
{
    int head_gt[10];
    const int nhead = n < 10 ? (int)n : 10;
    auto argmax = [](const float* v, int m) { int b = 0; for (int i = 1; i < m; ++i) if (v[i] > v[b]) b = i; return b; };
    for (int i = 0; i < nhead; ++i) head_gt[i] = argmax(gt_prob + (size_t)i * 21, 21);
    for (int64_t j = 0; j < n; ++j) {
        const float* gp = gt_prob + j * 21; const float* zp = zy_prob + j * 3;
        const int gt = argmax(gp, 21), zyo = argmax(zp, 3);
        if (gt >= 10) continue;
        const char sref = (char)refbase[j];
        const char* label = kGt[gt];
        const char* zy = kZy[zyo];
        const float* cov = cov8 + j * 8;
        float neg = 0.f; for (int k = 0; k < 8; ++k) if (cov[k] < 0.f) neg += cov[k];
        const float depth = -1.0f * neg;
        char alt[4]; int na = 0;
        for (int k = 0; k < 2; ++k) if (label[k] != sref) alt[na++] = label[k];
        alt[na] = 0;
        float support = 0.f; bool bad = false;
        for (int k = 0; k < na; ++k) { const int b = base_idx(alt[k]); if (b < 0) { bad = true; break; } support += cov[b]; support += cov[b + 4]; }
        if (bad) continue;
        const float af = support / depth;
        const bool af_one = af > 1.0f;
        double gt_q, zy_q;
        if (!calc_score_fast(gp[gt], &gt_q)) continue;
        if (!calc_score_fast(zp[zyo], &zy_q)) continue;
        // min() of the ROUNDED values in Python; rounding is monotone, so the smaller unrounded value rounds to the minimum
        const double qual = gt_q < zy_q ? gt_q : zy_q;
        char* e = nullptr;
        if (na == 0) {
            if (zyo == 0) {
                const char a1[2] = {sref, 0};
                e = put_record(p, contig, clen, pos1[j], sref, a1, qual, "RefCall", zy, depth, af, af_one);
            } else {
                static const int tis_hom[4] = {0, 4, 7, 9};
                static const int tis_het[6] = {1, 2, 3, 5, 6, 8};
                const int* tis = zyo == 1 ? tis_hom : tis_het; const int nt = zyo == 1 ? 4 : 6;
                int max_ti = -1, max_v = -1; bool raised = false;
                for (int q = 0; q < nt; ++q) {
                    const int ti = tis[q];
                    if (zyo == 1 && kGt[ti][0] == sref) continue;
                    if (ti >= n) { raised = true; break; }
                    if (head_gt[ti] > max_v) { max_v = head_gt[ti]; max_ti = ti; }
                }
                if (raised) continue;
                char a1[2] = {0, 0};
                if (zyo == 1) a1[0] = kGt[max_ti][0]; else a1[0] = kGt[max_ti][0] == sref ? kGt[max_ti][1] : kGt[max_ti][0];
                e = put_record(p, contig, clen, pos1[j], sref, a1, zy_q, "PASS", zy, depth, af, af_one);
            }
        } else {
            char alts[8];
            if (na == 1) { alts[0] = alt[0]; alts[1] = 0; }
            else if (alt[0] == alt[1]) { alts[0] = alt[0]; alts[1] = 0; }
            else { alts[0] = alt[0]; alts[1] = ','; alts[2] = alt[1]; alts[3] = 0; }
            if (alts[1] == ',' && zyo != 2) zy = "1/2";
            e = put_record(p, contig, clen, pos1[j], sref, alts, zyo == 0 ? gt_q : qual, "PASS", zy, depth, af, af_one);
        }
        p = e;
    }
    return p;
}

}  // namespace

namespace {
constexpr size_t kMaxRecordBytes = 128;     // upper bound of one record's text without the contig name

// Runs nt workers; work(t, p) writes the t-th part at p (capacity need(t) bytes) and returns its end.  The parts are then
// concatenated into `out` by the workers themselves once every size is known (spin barrier): a single-threaded memcpy of
// ~64 bytes per record cost as much as formatting the records on two threads.  Part buffers persist in a pool (no page
// faults after the first call).  Returns the total size, negated if it does not fit.
struct PartBuf { char* data = nullptr; size_t cap = 0; };
template <class Need, class Work>
int64_t run_parts_parallel(int nt, Need&& need, Work&& work, char* out, int64_t out_capacity) {
    static std::mutex pool_mu;
    static std::vector<PartBuf> pool;
    std::vector<PartBuf> parts((size_t)nt);
    {
        std::lock_guard<std::mutex> lk(pool_mu);
        for (int t = 0; t < nt && !pool.empty(); ++t) { parts[(size_t)t] = pool.back(); pool.pop_back(); }
    }
    std::vector<int64_t> sizes((size_t)nt, 0);
    std::atomic<int> done{0};
    std::atomic<int> failed{0};
    auto body = [&](int t) {
        PartBuf pb = parts[(size_t)t];                     // worked on locally: the entries of `parts` share cache lines
        const size_t want = need(t);
        if (pb.cap < want) { free(pb.data); pb.data = (char*)malloc(want + want / 8 + 64); pb.cap = pb.data ? want + want / 8 + 64 : 0; }
        if (!pb.data) failed.store(1);
        else sizes[(size_t)t] = (int64_t)(work(t, pb.data) - pb.data);
        done.fetch_add(1, std::memory_order_acq_rel);
        while (done.load(std::memory_order_acquire) < nt) std::this_thread::yield();
        int64_t off = 0, total = 0;
        for (int k = 0; k < nt; ++k) { if (k < t) off += sizes[(size_t)k]; total += sizes[(size_t)k]; }
        if (out && total <= out_capacity && !failed.load() && pb.data) memcpy(out + off, pb.data, (size_t)sizes[(size_t)t]);
        parts[(size_t)t] = pb;
    };
    if (nt == 1) body(0);
    else {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(body, t);
        body(0);
        for (auto& x : th) x.join();
    }
    int64_t total = 0;
    for (auto v : sizes) total += v;
    const bool fits = out && total <= out_capacity && !failed.load();
    {
        std::lock_guard<std::mutex> lk(pool_mu);
        for (auto& pb : parts) { if (pool.size() < 256) pool.push_back(pb); else free(pb.data); }
    }
    if (failed.load()) return 0;
    return fits ? total : -total;
}
}  // namespace

extern "C" int64_t nsnp_vcf_format_contig(const char* contig, int64_t n, const int32_t* pos1, const uint8_t* refbase,
                                          const float* gt_prob, const float* zy_prob, const float* cov8, int64_t batch_size,
                                          int n_threads, char* out, int64_t out_capacity)
{
    if (!contig || n < 0 || batch_size <= 0 || (n > 0 && (!pos1 || !refbase || !gt_prob || !zy_prob || !cov8))) return 0;
    const size_t clen = strlen(contig);
    if (clen > 200) return 0;
    const int64_t n_batches = (n + batch_size - 1) / batch_size;
    int nt = n_threads < 1 ? 1 : n_threads;
    if ((int64_t)nt > n_batches) nt = (int)(n_batches > 0 ? n_batches : 1);
    // per-thread text buffers are kept between calls (fresh 40 MB allocations page-fault under the process-wide mm lock
    // and stop the threads from scaling); the pool is guarded for concurrent callers
    auto need = [&](int t) { return (size_t)((n_batches * (t + 1) / nt - n_batches * t / nt) * batch_size) * (kMaxRecordBytes + clen) + 64; };
    return run_parts_parallel(nt, need, [&](int t, char* p) {
        const int64_t b0 = n_batches * t / nt, b1 = n_batches * (t + 1) / nt;
        for (int64_t b = b0; b < b1; ++b) {
            const int64_t s = b * batch_size, m = (n - s) < batch_size ? (n - s) : batch_size;
            p = format_batch_fast(p, contig, clen, m, pos1 + s, refbase + s, gt_prob + s * 21, zy_prob + s * 3, cov8 + s * 8);
        }
        return p;
    }, out, out_capacity);
}


// ---- text from compact GPU records (record.cu) -------------------------------------------------------------------
namespace {

inline char* put_q100(char* p, long long q100, long long* int_part) {
    const long long ip = q100 / 100; const int f2 = (int)(q100 % 100);
    *int_part = ip;
    p = put_uint(p, (unsigned long long)ip);
    *p++ = '.';
    *p++ = (char)('0' + f2 / 10);
    if (f2 % 10) *p++ = (char)('0' + f2 % 10);
    return p;
}

// QUAL*100 of a record field; rounding ties flagged by the device are redone with the libc path
inline bool rec_q100(float p, int32_t q_dev, bool tie, long long* q100) {
    if (!tie) { *q100 = q_dev; return true; }
    double v;
    if (!calc_score(p, &v)) return false;                  // rounded to 2 decimals by calc_score
    *q100 = (long long)floor(v * 100.0 + 0.5);
    return true;
}

// One record from already-decided fields.  Short strings are copied as fixed-size words and the pointer advanced by
// their length (alt <= 3, filter <= 7, zy = 3 characters; the caller's buffer has slack), and the integer part of QUAL,
// which is printed twice (QUAL and GQ), is converted once.
inline char* put_record_q(char* p, const char* contig, size_t clen, long long pos, char ref, const char* alt, long long q100,
                          const char* filter, const char* zy, long long depth, int32_t af_q)
{
    memcpy(p, contig, clen); p += clen; *p++ = '\t';
    p = put_u32(p, (uint32_t)pos);
    memcpy(p, "\t.\t", 3); p[3] = ref; p[4] = '\t'; p += 5;
    { char a4[4] = {0, 0, 0, 0}; int na = 0; while (alt[na]) { a4[na] = alt[na]; ++na; } memcpy(p, a4, 4); p += na; }
    *p++ = '\t';
    const uint32_t ip = (uint32_t)(q100 / 100); const uint32_t f2 = (uint32_t)(q100 - (long long)ip * 100);
    char ipd[12]; const int ipn = (int)(put_u32(ipd, ip) - ipd);
    memcpy(p, ipd, 12); p += ipn;
    *p++ = '.';
    *p++ = (char)('0' + f2 / 10);
    if (f2 % 10) *p++ = (char)('0' + f2 % 10);
    *p++ = '\t';
    if (filter[0] == 'P') { memcpy(p, "PASS", 4); p += 4; } else { memcpy(p, "RefCall", 7); p += 7; }
    memcpy(p, "\t.\tGT:GQ:DP:AF\t", 15); p += 15;
    memcpy(p, zy, 3); p[3] = ':'; p += 4;
    memcpy(p, ipd, 12); p += ipn;
    *p++ = ':';
    if (depth < 0) { *p++ = '-'; p = put_uint(p, (unsigned long long)(-depth)); } else p = put_u32(p, (uint32_t)depth);
    *p++ = ':';
    if (af_q == NSNP_AF_ONE) { memcpy(p, "1.000000", 8); p += 8; }
    else if (af_q == NSNP_AF_NAN) { memcpy(p, "nan", 3); p += 3; }
    else if (af_q == NSNP_AF_NEG_INF) { memcpy(p, "-inf", 4); p += 4; }
    else {
        if (af_q <= -3) { *p++ = '-'; af_q = -(af_q + 3); }
        const uint32_t u = (uint32_t)af_q, ipa = u / 1000000u, f = u - ipa * 1000000u;
        p = put_u32(p, ipa);
        *p++ = '.';
        const uint32_t f01 = f / 10000u, f2345 = f - f01 * 10000u, f23 = f2345 / 100u, f45 = f2345 - f23 * 100u;
        memcpy(p, kDigits2 + 2 * f01, 2); memcpy(p + 2, kDigits2 + 2 * f23, 2); memcpy(p + 4, kDigits2 + 2 * f45, 2); p += 6;
    }
    *p++ = '\n';
    return p;
}

char* format_batch_records(char* p, const char* contig, size_t clen, int64_t n, const nsnp_site_record_t* rec) {
    for (int64_t j = 0; j < n; ++j) {
        const nsnp_site_record_t& r = rec[j];
        const int gt = r.gt, zyo = r.zy;
        if (gt >= 10) continue;                                               // predict.py:68
        if (r.flags & NSNP_REC_DROP) continue;                                // calculate_score raised
        const char sref = (char)r.ref;
        const char* label = kGt[gt];
        const char* zy = kZy[zyo];
        char alt[4]; int na = 0;
        for (int k = 0; k < 2; ++k) if (label[k] != sref) alt[na++] = label[k];
        alt[na] = 0;
        long long gt_q, zy_q;
        if (!rec_q100(r.p_gt, r.q100_gt, (r.flags & NSNP_REC_TIE_GT) != 0, &gt_q)) continue;
        if (!rec_q100(r.p_zy, r.q100_zy, (r.flags & NSNP_REC_TIE_ZY) != 0, &zy_q)) continue;
        const long long qual = gt_q < zy_q ? gt_q : zy_q;
        char* e = nullptr;
        if (na == 0) {
            if (zyo == 0) {
                const char a1[2] = {sref, 0};
                e = put_record_q(p, contig, clen, r.pos1, sref, a1, qual, "RefCall", zy, r.depth, r.af_q);
            } else {
                static const int tis_hom[4] = {0, 4, 7, 9};
                static const int tis_het[6] = {1, 2, 3, 5, 6, 8};
                const int* tis = zyo == 1 ? tis_hom : tis_het; const int nt = zyo == 1 ? 4 : 6;
                int max_ti = -1, max_v = -1; bool raised = false;
                for (int q = 0; q < nt; ++q) {
                    const int ti = tis[q];
                    if (zyo == 1 && kGt[ti][0] == sref) continue;
                    if (ti >= n) { raised = true; break; }                    // IndexError on the batch argmax array
                    const int v = rec[ti].gt;                                 // gt_output[ti]: the batch array indexed by a class index
                    if (v > max_v) { max_v = v; max_ti = ti; }
                }
                if (raised) continue;
                char a1[2] = {0, 0};
                if (zyo == 1) a1[0] = kGt[max_ti][0]; else a1[0] = kGt[max_ti][0] == sref ? kGt[max_ti][1] : kGt[max_ti][0];
                e = put_record_q(p, contig, clen, r.pos1, sref, a1, zy_q, "PASS", zy, r.depth, r.af_q);
            }
        } else {
            char alts[8];
            if (na == 1) { alts[0] = alt[0]; alts[1] = 0; }
            else if (alt[0] == alt[1]) { alts[0] = alt[0]; alts[1] = 0; }
            else { alts[0] = alt[0]; alts[1] = ','; alts[2] = alt[1]; alts[3] = 0; }
            if (alts[1] == ',' && zyo != 2) zy = "1/2";
            e = put_record_q(p, contig, clen, r.pos1, sref, alts, zyo == 0 ? gt_q : qual, "PASS", zy, r.depth, r.af_q);
        }
        p = e;
    }
    return p;
}

}  // namespace

extern "C" int64_t nsnp_vcf_format_contig_records(const char* contig, int64_t n, const nsnp_site_record_t* rec, int64_t batch_size,
                                                  int n_threads, char* out, int64_t out_capacity)
{
    if (!contig || n < 0 || batch_size <= 0 || (n > 0 && !rec)) return 0;
    const size_t clen = strlen(contig);
    if (clen > 200) return 0;
    const int64_t n_batches = (n + batch_size - 1) / batch_size;
    int nt = n_threads < 1 ? 1 : n_threads;
    if ((int64_t)nt > n_batches) nt = (int)(n_batches > 0 ? n_batches : 1);
    auto need = [&](int t) { return (size_t)((n_batches * (t + 1) / nt - n_batches * t / nt) * batch_size) * (kMaxRecordBytes + clen) + 64; };
    return run_parts_parallel(nt, need, [&](int t, char* p) {
        const int64_t b0 = n_batches * t / nt, b1 = n_batches * (t + 1) / nt;
        for (int64_t b = b0; b < b1; ++b) {
            const int64_t s = b * batch_size, m = (n - s) < batch_size ? (n - s) : batch_size;
            p = format_batch_records(p, contig, clen, m, rec + s);
        }
        return p;
    }, out, out_capacity);
}
