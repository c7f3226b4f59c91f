// Synthetic ONT-like read generator: integer-only arithmetic so the host (gcc)
// and device (nvcc) builds produce bit-identical reads.  All randomness is counter based: every value
// is a hash of (seed, read index, counter), so any read can be generated independently.
//
// CIGARs stay inside the unambiguous subset of SURVEY appendix B.4:
//     [S] M { (I|D) M }* [S]      (M may be emitted as =/X with use_eqx)
#pragma once
#include <stdint.h>
#include "../../include/nanosnp_b200.h"

#ifdef __CUDACC__
#define NSNP_HD __host__ __device__ __forceinline__
#else
#define NSNP_HD static inline
#endif

NSNP_HD uint64_t nsnp_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
NSNP_HD uint64_t nsnp_hash3(uint64_t seed, uint64_t a, uint64_t b) {
    return nsnp_mix64(nsnp_mix64(seed ^ nsnp_mix64(a)) + b);
}

// ---- reference and planted variants -----------------------------------------------------------
NSNP_HD int nsnp_ref_code(const nsnp_synth_cfg_t* c, int64_t p) {   // 0..3
    return (int)(nsnp_hash3(c->seed_ref, 0x5EF, (uint64_t)p) & 3);
}
NSNP_HD uint8_t nsnp_ref_char(const nsnp_synth_cfg_t* c, int64_t p) {
    const char up[4] = {'A', 'C', 'G', 'T'};
    uint8_t ch = (uint8_t)up[nsnp_ref_code(c, p)];
    if (c->ref_n_period > 0 && (p % c->ref_n_period) >= c->ref_n_period - c->ref_n_len) {
        // mix upper and lower case N plus one IUPAC code to exercise evc_base_from()
        int k = (int)(p % 3);
        ch = k == 0 ? 'N' : (k == 1 ? 'n' : 'R');
    } else if (c->ref_lower_period > 0 && (p % c->ref_lower_period) < c->ref_lower_len) {
        ch = (uint8_t)(ch + 32);
    }
    return ch;
}
// planted SNP: returns alt code (0..3) and sets *hom, or -1
NSNP_HD int nsnp_variant(const nsnp_synth_cfg_t* c, int64_t p, int* hom) {
    uint64_t h = nsnp_hash3(c->seed_var, 0x7A8, (uint64_t)p);
    if ((uint32_t)h >= c->snp_thr) return -1;
    int r = nsnp_ref_code(c, p);
    int alt = (r + 1 + (int)((h >> 32) % 3)) & 3;
    *hom = ((h >> 40) % 3) == 0;     // het : hom = 2 : 1
    return alt;
}

// ---- per-read header ------------------------------------------------------------------------------
struct nsnp_read_hdr {
    int32_t  pos;      // 0-based start
    int32_t  span;     // reference span (0 => read carries only a 1M placeholder and is flagged unmapped)
    uint16_t flag;
    uint8_t  mapq;
    uint8_t  hap;
    int32_t  clip5, clip3;
};

NSNP_HD nsnp_read_hdr nsnp_read_header(const nsnp_synth_cfg_t* c, int64_t r) {
    nsnp_read_hdr h;
    uint64_t a = nsnp_hash3(c->seed_reads, (uint64_t)r, 1);
    uint64_t b = nsnp_hash3(c->seed_reads, (uint64_t)r, 2);
    uint64_t d = nsnp_hash3(c->seed_reads, (uint64_t)r, 3);
    // stratified-jitter start: sorted by construction
    const uint64_t L = (uint64_t)c->contig_len;
    const uint64_t u16 = a & 0xFFFF;                                    // jitter in [0,1) * 2^16
    // pos = floor(r * L / n_reads) + jitter, jitter < ceil(L / n_reads): non-decreasing in r
    const uint64_t n = (uint64_t)c->n_reads;
    const uint64_t stride = (L + n - 1) / n;
    uint64_t pos = ((uint64_t)r * L) / n + ((u16 * stride) >> 16);
    if (pos >= L) pos = L - 1;
    // span from the quantile table with linear interpolation
    const uint32_t q = (uint32_t)(a >> 16) & 0x3FF;                      // 0..1023
    const uint32_t fr = (uint32_t)(a >> 26) & 0xFFFF;
    const int64_t q0 = c->len_quantiles[q], q1 = c->len_quantiles[q + 1];
    int64_t span = q0 + (((q1 - q0) * (int64_t)fr) >> 16);
    if ((int64_t)pos + span > (int64_t)L) span = (int64_t)L - (int64_t)pos;
    h.flag = (b & 1) ? 16 : 0;
    h.hap = (uint8_t)((b >> 1) & 1);
    h.mapq = 60;
    const uint32_t u = (uint32_t)(b >> 32);
    if (u < c->lowmapq_thr) h.mapq = (uint8_t)((b >> 8) % 20);
    const uint32_t v = (uint32_t)d;
    if (v < c->secondary_thr) h.flag |= 256;
    else if (v - c->secondary_thr < c->supp_thr && v >= c->secondary_thr) h.flag |= 2048;
    // coverage gaps: truncate reads running into a gap, drop reads starting inside one
    if (c->gap_period > 0) {
        const int64_t gp = c->gap_period, gl = c->gap_len;
        const int64_t off = (int64_t)pos % gp;          // gap occupies [gp-gl, gp) of every period
        if (off >= gp - gl) span = 0;
        else { const int64_t room = (gp - gl) - off; if (span > room) span = room; }
    }
    if (span < c->len_min) { span = 0; h.flag |= 4; }
    h.pos = (int32_t)pos;
    h.span = (int32_t)span;
    h.clip5 = h.clip3 = 0;
    if ((uint32_t)(d >> 32) < c->softclip_thr) {
        h.clip5 = 1 + (int32_t)((d >> 8) % 40);
        h.clip3 = (int32_t)((d >> 16) % 40);
    }
    return h;
}

NSNP_HD int nsnp_cdf_sample(const uint32_t* cdf, int n, uint32_t u) {   // smallest k with u < cdf[k], else n-1
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (u < cdf[mid]) hi = mid; else lo = mid + 1; }
    return lo;
}

// Generic walker: calls emit_op(op, len) for each CIGAR op and emit_base(code, isN) for each SEQ base.
template <class OpF, class BaseF>
NSNP_HD void nsnp_walk_read(const nsnp_synth_cfg_t* c, int64_t r, const nsnp_read_hdr& h, OpF emit_op, BaseF emit_base) {
    uint64_t ctr = 16;
    int64_t qi = 0;
    auto rnd = [&]() { return nsnp_hash3(c->seed_reads, (uint64_t)r, ctr++); };
    auto rand_base = [&]() { uint64_t x = rnd(); emit_base((int)(x & 3), (uint32_t)(x >> 32) < c->nbase_thr); ++qi; };
    if (h.span <= 0) {                  // filtered placeholder: 1M, one base
        emit_op(0, 1); emit_base(0, false); return;
    }
    if (h.clip5 > 0) { emit_op(4, h.clip5); for (int i = 0; i < h.clip5; ++i) rand_base(); }
    int64_t p = h.pos;
    int64_t remaining = h.span;
    const uint32_t indel_thr_sum = c->ins_thr + c->del_thr;
    while (remaining > 0) {
        uint64_t x = rnd();
        int64_t m = 1 + nsnp_cdf_sample(c->mrun_cdf, 256, (uint32_t)x);
        if (indel_thr_sum == 0 || m > remaining) m = remaining;
        // aligned run: emitted as M, or as =/X runs
        int64_t run_start = 0; int run_kind = -1;
        for (int64_t j = 0; j < m; ++j, ++p) {
            uint64_t y = rnd();
            int b = nsnp_ref_code(c, p);
            int hom = 0; const int alt = nsnp_variant(c, p, &hom);
            if (alt >= 0 && (hom || h.hap)) b = alt;
            if ((uint32_t)y < c->sub_thr) b = (b + 1 + (int)((y >> 32) % 3)) & 3;
            const bool isN = (uint32_t)(y >> 34) < (c->nbase_thr >> 2);
            emit_base(b, isN); ++qi;
            if (c->use_eqx) {
                // '=' iff read base equals the (ACGT) reference letter; N bases and non-ACGT reference count as X
                const uint8_t rc = nsnp_ref_char(c, p);
                const bool acgt = rc == 'A' || rc == 'C' || rc == 'G' || rc == 'T' || rc == 'a' || rc == 'c' || rc == 'g' || rc == 't';
                const int kind = (!isN && acgt && b == nsnp_ref_code(c, p)) ? 7 : 8;
                if (kind != run_kind) { if (run_kind >= 0) emit_op(run_kind, (int)(j - run_start)); run_kind = kind; run_start = j; }
            }
        }
        if (c->use_eqx) emit_op(run_kind, (int)(m - run_start)); else emit_op(0, (int)m);
        remaining -= m;
        if (remaining <= 0) break;
        // an indel follows; the read must still end with an aligned run
        uint64_t z = rnd();
        int len = 1 + nsnp_cdf_sample(c->indel_cdf, 60, (uint32_t)z);
        if ((uint32_t)(z >> 32) < c->long_indel_thr) len = 61 + (int)((z >> 20) % 30);
        const bool is_ins = ((uint32_t)(rnd() >> 16) % indel_thr_sum) < c->ins_thr;
        if (is_ins) {
            emit_op(1, len);
            for (int i = 0; i < len; ++i) rand_base();
        } else {
            if (len > remaining - 1) len = (int)(remaining - 1);
            if (len > 0) { emit_op(2, len); p += len; remaining -= len; }
            else { emit_op(1, 1); rand_base(); }      // no room for a deletion: make it a 1-base insertion
        }
    }
    if (h.clip3 > 0) { emit_op(4, h.clip3); for (int i = 0; i < h.clip3; ++i) rand_base(); }
}
