// Host-side BAM record decoder: uncompressed BAM bytes -> flat packed read arrays (struct nsnp_reads).
// This is the "host decodes the BAM into flat packed arrays" step of the north star; it replaces what samtools does
// before mpileup (make_predict_data.sh:151).  BGZF inflation is done by the caller (zlib); this file only parses.
// BAM layout per the SAM/BAM specification section 4.2; CIGARs with more than 65535 ops are taken from the CG:B,I tag.
#include "common.cuh"

namespace {

inline int32_t rd_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint32_t rd_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd_u16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

struct Rec {
    int32_t ref_id, pos; uint32_t l_read_name, mapq, n_cigar, flag, l_seq;
    const uint8_t* cigar; const uint8_t* seq; const uint8_t* tags; const uint8_t* end;
};

// returns bytes consumed, 0 at a clean end, -1 on corruption
inline int64_t parse_rec(const uint8_t* p, const uint8_t* lim, Rec* r) {
    if (p == lim) return 0;
    if (lim - p < 36) return -1;
    const int64_t bs = rd_i32(p);
    if (bs < 32 || p + 4 + bs > lim) return -1;
    r->ref_id = rd_i32(p + 4); r->pos = rd_i32(p + 8);
    r->l_read_name = p[12]; r->mapq = p[13];
    r->n_cigar = rd_u16(p + 16); r->flag = rd_u16(p + 18); r->l_seq = rd_u32(p + 20);
    const uint8_t* q = p + 36 + r->l_read_name;
    r->cigar = q; q += 4ull * r->n_cigar;
    r->seq = q; q += (r->l_seq + 1) / 2;
    q += r->l_seq;                               // qualities: never read (--min-BQ 0)
    r->tags = q; r->end = p + 4 + bs;
    if (q > r->end) return -1;
    return 4 + bs;
}

// real CIGAR of a record: the CG:B,I tag when the in-record CIGAR is the <read length>S<ref length>N placeholder
inline void real_cigar(const Rec& r, const uint8_t** cig, uint32_t* n) {
    *cig = r.cigar; *n = r.n_cigar;
    if (r.n_cigar != 2) return;
    const uint32_t c0 = rd_u32(r.cigar), c1 = rd_u32(r.cigar + 4);
    if ((c0 & 15) != 4 || (c0 >> 4) != r.l_seq || (c1 & 15) != 3) return;
    const uint8_t* t = r.tags;
    while (t + 3 <= r.end) {
        const char a = (char)t[0], b = (char)t[1], ty = (char)t[2];
        t += 3;
        size_t sz = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'Z': case 'H': { const uint8_t* e = t; while (e < r.end && *e) ++e; sz = (size_t)(e - t) + 1; break; }
            case 'B': {
                if (t + 5 > r.end) return;
                const char sub = (char)t[0]; const uint32_t cnt = rd_u32(t + 1);
                const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                if (a == 'C' && b == 'G' && sub == 'I' && t + 5 + 4ull * cnt <= r.end) { *cig = t + 5; *n = cnt; return; }
                sz = 5 + es * cnt; break;
            }
            default: return;
        }
        t += sz;
    }
}

// copies a record's CIGAR, merging adjacent ops of the same type; returns the number of ops written
inline uint32_t copy_cigar_merged(uint32_t* dst, const uint8_t* src, uint32_t n) {
    uint32_t w = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t c = rd_u32(src + 4ull * k);
        if (w && (dst[w - 1] & 15u) == (c & 15u)) dst[w - 1] += c & ~15u;
        else dst[w++] = c;
    }
    return w;
}

}  // namespace

extern "C" {

// Pass 1: counts for the records of reference `ref_id` (file order).  bases_padded: every read starts on a 16-base
// boundary of seq2.  Returns the number of reads, or -1 on a malformed file.
int64_t nsnp_bam_count(const uint8_t* data, int64_t n_bytes, int64_t first_record_offset, int32_t ref_id,
                       int64_t* n_cigar, int64_t* n_bases_padded)
{
    if (!data || first_record_offset < 0 || first_record_offset > n_bytes) return -1;
    const uint8_t* p = data + first_record_offset; const uint8_t* lim = data + n_bytes;
    int64_t reads = 0, ops = 0, bases = 0;
    Rec r;
    for (;;) {
        const int64_t used = parse_rec(p, lim, &r);
        if (used == 0) break;
        if (used < 0) return -1;
        if (r.ref_id == ref_id) {
            const uint8_t* cg; uint32_t nc; real_cigar(r, &cg, &nc);
            ++reads; ops += nc; bases += ((int64_t)r.l_seq + 15) / 16 * 16;
        }
        p += used;
    }
    if (n_cigar) *n_cigar = ops;
    if (n_bases_padded) *n_bases_padded = bases;
    return reads;
}

// Pass 2: fills the arrays sized by pass 1 (seq2 / nmask must be zero-initialised and hold n_bases_padded + 64 bases).
int64_t nsnp_bam_fill(const uint8_t* data, int64_t n_bytes, int64_t first_record_offset, int32_t ref_id,
                      int32_t* pos, uint16_t* flag, uint8_t* mapq, int64_t* cigar_off, uint32_t* cigar, int64_t* seq_off,
                      uint8_t* seq2, uint8_t* nmask)
{
    if (!data || !pos || !flag || !mapq || !cigar_off || !cigar || !seq_off || !seq2) return -1;
    const uint8_t* p = data + first_record_offset; const uint8_t* lim = data + n_bytes;
    // 4-bit BAM base code "=ACMGRSVTWYHKDBN" -> 2-bit code / N flag
    static const int8_t code2[16] = {-1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1};
    int64_t i = 0, oi = 0, bi = 0;
    Rec r;
    cigar_off[0] = 0;
    for (;;) {
        const int64_t used = parse_rec(p, lim, &r);
        if (used == 0) break;
        if (used < 0) return -1;
        if (r.ref_id == ref_id) {
            const uint8_t* cg; uint32_t nc; real_cigar(r, &cg, &nc);
            pos[i] = r.pos; flag[i] = (uint16_t)r.flag; mapq[i] = (uint8_t)r.mapq;
            oi += copy_cigar_merged(cigar + oi, cg, nc);         // "1D2D" -> "3D": one indel, as htslib reports it
            cigar_off[i + 1] = oi;
            seq_off[i] = bi;
            for (uint32_t k = 0; k < r.l_seq; ++k) {
                const int c4 = (r.seq[k >> 1] >> ((k & 1) ? 0 : 4)) & 15;
                const int c2 = code2[c4];
                const int64_t b = bi + k;
                if (c2 >= 0) seq2[b >> 2] |= (uint8_t)(c2 << (2 * (b & 3)));
                else if (nmask) nmask[b >> 3] |= (uint8_t)(1u << (b & 7));
            }
            bi += ((int64_t)r.l_seq + 15) / 16 * 16;
            ++i;
        }
        p += used;
    }
    return i;
}

}  // extern "C"
