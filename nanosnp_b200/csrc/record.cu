// s2, numeric half of the VCF record logic on the GPU (predict.py:54-88): per site argmax / max probability of both
// heads, DP, AF and the two QUAL values, packed into a 32-byte record.  The host then only assembles text
// (nsnp_vcf_format_contig_records), so the device->host copy shrinks from 133 to 32 bytes per site and the host does
// no log() and no 21-way argmax per record.
//
// Exactness: all float32 steps (1-p, (1-p)/p, the DP sum, the AF quotient) are IEEE operations identical to NumPy's;
// the log-odds is evaluated in double like Python's math.log.  A double log() on the device may differ from glibc's in
// the last ulp, which can only change the 2-decimal rounding when the value sits within ~1e-12 of a rounding tie:
// anything within 1e-6 of a tie is flagged and recomputed on the host from the probability carried in the record.
#include <math.h>
#include "common.cuh"

namespace nsnp {
namespace {

__constant__ char kGtLabel[10][2] = {{'A', 'A'}, {'A', 'C'}, {'A', 'G'}, {'A', 'T'}, {'C', 'C'}, {'C', 'G'}, {'C', 'T'}, {'G', 'G'}, {'G', 'T'}, {'T', 'T'}};

__device__ __forceinline__ int base_index(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// calculate_score (predict.py:31-34) up to, not including, round(): returns false when Python raises (log of <= 0)
__device__ __forceinline__ bool score_q100(float p, double kScale, int32_t* q100, bool* tie) {
    const float a = __fsub_rn(1.0f, p);
    const float r = __fdiv_rn(a, p);
    const double x = (double)r;
    if (!(x > 0.0)) return false;
    double t = __dadd_rn(__dmul_rn(kScale, log(x)), 10.0);
    if (!(t > 0.0)) t = 0.0;
    const double h = __dmul_rn(t, 100.0);
    const double fl = floor(h), fr = h - fl;
    *tie = fabs(fr - 0.5) < 1e-6 || !(h < 2.0e9);
    *q100 = (int32_t)fl + (fr > 0.5 ? 1 : 0);
    return true;
}

__global__ void __launch_bounds__(256) site_record_kernel(const float* __restrict__ gt, const float* __restrict__ zy,
                                                         const int32_t* __restrict__ x, const uint8_t* __restrict__ refbase,
                                                         const int32_t* __restrict__ pos0, int64_t n, const int32_t* __restrict__ n_dev,
                                                         nsnp_site_record_t* __restrict__ rec, double kScale,
                                                         const int32_t* __restrict__ counts, int64_t region_start, const uint8_t* __restrict__ ref)
{
    // counts != nullptr: the centre row and the reference base come straight from the count tensor / the contig (no window tensor)
    if (n_dev) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const float* gp = gt + j * NSNP_GT_CLASSES; const float* zp = zy + j * NSNP_ZY_CLASSES;
        int gi = 0; float gm = gp[0];
#pragma unroll
        for (int k = 1; k < NSNP_GT_CLASSES; ++k) { const float v = gp[k]; if (v > gm) { gm = v; gi = k; } }       // np.argmax: first maximum
        int zi = 0; float zm = zp[0];
        if (zp[1] > zm) { zm = zp[1]; zi = 1; }
        if (zp[2] > zm) { zm = zp[2]; zi = 2; }
        nsnp_site_record_t r;
        uint8_t rb;
        if (counts) { rb = ref[pos0[j]]; if (rb >= 'a' && rb <= 'z') rb = (uint8_t)(rb - 32); } else rb = refbase[j];
        r.gt = (uint8_t)gi; r.zy = (uint8_t)zi; r.flags = 0; r.ref = rb;
        r.pos1 = pos0[j] + 1; r.p_gt = gm; r.p_zy = zm; r.q100_gt = 0; r.q100_zy = 0; r.depth = 0; r.af_q = 0;
        // centre row, channels [A C G T a c g t] (predict.py:63)
        const int32_t* row = counts ? counts + ((int64_t)pos0[j] - region_start) * NSNP_CHANNELS : x + (j * NSNP_WINDOW + NSNP_FLANK) * NSNP_CHANNELS;
        float cov[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cov[k] = (float)row[k]; cov[4 + k] = (float)row[9 + k]; }
        float neg = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (cov[k] < 0.f) neg = __fadd_rn(neg, cov[k]);
        const float depth = -1.0f * neg;                                  // predict.py:76
        r.depth = (int32_t)depth;
        if (gi < 10) {
            float support = 0.f;
            bool bad = false;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const char c = kGtLabel[gi][k];
                if (c != (char)r.ref) { const int b = base_index(c); if (b < 0) bad = true; else { support = __fadd_rn(support, cov[b]); support = __fadd_rn(support, cov[b + 4]); } }
            }
            if (bad) r.flags |= NSNP_REC_DROP;
            const float af = __fdiv_rn(support, depth);                  // float32 quotient: 0/0 -> nan, x/0 -> inf
            if (af > 1.0f) r.af_q = NSNP_AF_ONE;                          // predict.py:83-84
            else if (af != af) r.af_q = NSNP_AF_NAN;
            else if (af == -INFINITY) r.af_q = NSNP_AF_NEG_INF;
            else {
                const bool neg = signbit(af);
                const double m = fabs((double)af) * 1.0e6;                // exact: 24-bit significand x 20-bit integer
                double q = floor(m); const double fr = m - q;
                if (fr > 0.5 || (fr == 0.5 && fmod(q, 2.0) == 1.0)) q += 1.0;   // printf('%f') rounds the exact value, ties to even
                if (q > 2.0e9) q = 2.0e9;
                r.af_q = neg ? -(int32_t)q - 3 : (int32_t)q;
            }
            bool tie = false;
            if (!score_q100(gm, kScale, &r.q100_gt, &tie)) r.flags |= NSNP_REC_DROP; else if (tie) r.flags |= NSNP_REC_TIE_GT;
            tie = false;
            if (!score_q100(zm, kScale, &r.q100_zy, &tie)) r.flags |= NSNP_REC_DROP; else if (tie) r.flags |= NSNP_REC_TIE_ZY;
        }
        rec[j] = r;
    }
}

}  // namespace
}  // namespace nsnp

using namespace nsnp;

extern "C" int nsnp_site_records(const float* gt_prob_dev, const float* zy_prob_dev, const int32_t* x_i32_dev, const uint8_t* refbase_dev,
                                 const int32_t* pos_dev, int64_t n, const int32_t* n_dev, nsnp_site_record_t* rec_dev, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return NSNP_OK;
    if (!gt_prob_dev || !zy_prob_dev || !x_i32_dev || !refbase_dev || !pos_dev || !rec_dev || n < 0)
        return set_error(NSNP_E_INVALID, "nsnp_site_records: null argument");
    if (nsnp_device_count() == 0) return set_error(NSNP_E_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    int64_t blocks = (n + 255) / 256; if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    ProfScope prof(NSNP_PROF_RECORDS, stream);
    const double kScale = -10.0 * (1.0 / log(10.0));          // -10 * log(e, 10) exactly as the host formatter (vcf.cu) forms it
    site_record_kernel<<<(int)blocks, 256, 0, stream>>>(gt_prob_dev, zy_prob_dev, x_i32_dev, refbase_dev, pos_dev, n, n_dev, rec_dev, kScale, nullptr, 0, nullptr);
    return cuda_status("site_record_kernel");
}

extern "C" int nsnp_site_records_sites(const float* gt_prob_dev, const float* zy_prob_dev, const int32_t* counts_dev, int64_t region_start,
                                       const uint8_t* ref_dev, const int32_t* pos_dev, int64_t n, const int32_t* n_dev, nsnp_site_record_t* rec_dev,
                                       void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return NSNP_OK;
    if (!gt_prob_dev || !zy_prob_dev || !counts_dev || !ref_dev || !pos_dev || !rec_dev || n < 0)
        return set_error(NSNP_E_INVALID, "nsnp_site_records_sites: null argument");
    if (nsnp_device_count() == 0) return set_error(NSNP_E_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    int64_t blocks = (n + 255) / 256; if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    ProfScope prof(NSNP_PROF_RECORDS, stream);
    const double kScale = -10.0 * (1.0 / log(10.0));
    site_record_kernel<<<(int)blocks, 256, 0, stream>>>(gt_prob_dev, zy_prob_dev, nullptr, nullptr, pos_dev, n, n_dev, rec_dev, kScale, counts_dev, region_start, ref_dev);
    return cuda_status("site_record_kernel");
}
