// Synthetic read generator entry points (host loops + one-thread-per-read CUDA kernels).
// Not part of the measured path: it only manufactures the inputs of bench.py and the tests.
#include "common.cuh"
#include "synth_core.h"

namespace {

struct CountOut { int32_t n_ops = 0; int32_t n_query = 0; };

template <class T>
NSNP_HD void synth_count_one(const nsnp_synth_cfg_t* c, int64_t r, int32_t* pos, uint16_t* flag, uint8_t* mapq,
                             int32_t* n_ops, int32_t* n_query) {
    const nsnp_read_hdr h = nsnp_read_header(c, r);
    int32_t ops = 0, q = 0;
    nsnp_walk_read(c, r, h, [&](int, int) { ++ops; }, [&](int, bool) { ++q; });
    pos[r] = h.pos; flag[r] = h.flag; mapq[r] = h.mapq; n_ops[r] = ops; n_query[r] = q;
}

NSNP_HD void synth_fill_one(const nsnp_synth_cfg_t* c, int64_t r, const int64_t* cigar_off, const int64_t* seq_off,
                            uint32_t* cigar, uint8_t* seq2, uint8_t* nmask) {
    const nsnp_read_hdr h = nsnp_read_header(c, r);
    int64_t oi = cigar_off[r];
    int64_t bi = seq_off[r];           // multiple of 16 by contract of the generator: no byte is shared by two reads
    uint32_t acc2 = 0, accn = 0; int fill = 0;
    nsnp_walk_read(c, r, h,
        [&](int op, int len) { cigar[oi++] = ((uint32_t)len << 4) | (uint32_t)op; },
        [&](int code, bool isN) {
            acc2 |= (uint32_t)code << (2 * (fill & 3));
            accn |= (uint32_t)(isN ? 1 : 0) << (fill & 7);
            ++fill;
            if ((fill & 3) == 0) { seq2[(bi + fill - 4) >> 2] = (uint8_t)acc2; acc2 = 0; }
            if ((fill & 7) == 0) { if (nmask) nmask[(bi + fill - 8) >> 3] = (uint8_t)accn; accn = 0; }
        });
    if (fill & 3) seq2[(bi + (fill & ~3)) >> 2] = (uint8_t)acc2;
    if ((fill & 7) && nmask) nmask[(bi + (fill & ~7)) >> 3] = (uint8_t)accn;
}

__global__ void synth_ref_kernel(nsnp_synth_cfg_t c, uint8_t* ref) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < c.contig_len; p += (int64_t)gridDim.x * blockDim.x)
        ref[p] = nsnp_ref_char(&c, p);
}
__global__ void synth_count_kernel(nsnp_synth_cfg_t c, int32_t* pos, uint16_t* flag, uint8_t* mapq, int32_t* n_ops, int32_t* n_query) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < c.n_reads) synth_count_one<int>(&c, r, pos, flag, mapq, n_ops, n_query);
}
__global__ void synth_fill_kernel(nsnp_synth_cfg_t c, const int64_t* cigar_off, const int64_t* seq_off, uint32_t* cigar, uint8_t* seq2, uint8_t* nmask) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < c.n_reads) synth_fill_one(&c, r, cigar_off, seq_off, cigar, seq2, nmask);
}

int check_cfg(const nsnp_synth_cfg_t* c) {
    if (!c || c->contig_len <= 0 || c->n_reads < 0 || !c->len_quantiles || !c->mrun_cdf || !c->indel_cdf)
        return nsnp::set_error(NSNP_E_INVALID, "nsnp_synth: bad config");
    return NSNP_OK;
}

}  // namespace

extern "C" {

int nsnp_synth_ref_host(const nsnp_synth_cfg_t* cfg, uint8_t* ref_out) {
    if (int e = check_cfg(cfg)) return e;
    for (int64_t p = 0; p < cfg->contig_len; ++p) ref_out[p] = nsnp_ref_char(cfg, p);
    return NSNP_OK;
}
int nsnp_synth_count_host(const nsnp_synth_cfg_t* cfg, int32_t* pos, uint16_t* flag, uint8_t* mapq, int32_t* n_ops, int32_t* n_query) {
    if (int e = check_cfg(cfg)) return e;
    for (int64_t r = 0; r < cfg->n_reads; ++r) synth_count_one<int>(cfg, r, pos, flag, mapq, n_ops, n_query);
    return NSNP_OK;
}
int nsnp_synth_fill_host(const nsnp_synth_cfg_t* cfg, const int64_t* cigar_off, const int64_t* seq_off, uint32_t* cigar, uint8_t* seq2, uint8_t* nmask) {
    if (int e = check_cfg(cfg)) return e;
    for (int64_t r = 0; r < cfg->n_reads; ++r) synth_fill_one(cfg, r, cigar_off, seq_off, cigar, seq2, nmask);
    return NSNP_OK;
}
int nsnp_synth_ref_dev(const nsnp_synth_cfg_t* cfg, uint8_t* ref_dev, void* stream) {
    if (int e = check_cfg(cfg)) return e;
    synth_ref_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(*cfg, ref_dev);
    return nsnp::cuda_status("synth_ref_kernel");
}
int nsnp_synth_count_dev(const nsnp_synth_cfg_t* cfg, int32_t* pos, uint16_t* flag, uint8_t* mapq, int32_t* n_ops, int32_t* n_query, void* stream) {
    if (int e = check_cfg(cfg)) return e;
    if (cfg->n_reads == 0) return NSNP_OK;
    synth_count_kernel<<<(unsigned)((cfg->n_reads + 63) / 64), 64, 0, (cudaStream_t)stream>>>(*cfg, pos, flag, mapq, n_ops, n_query);
    return nsnp::cuda_status("synth_count_kernel");
}
int nsnp_synth_fill_dev(const nsnp_synth_cfg_t* cfg, const int64_t* cigar_off, const int64_t* seq_off, uint32_t* cigar, uint8_t* seq2, uint8_t* nmask, void* stream) {
    if (int e = check_cfg(cfg)) return e;
    if (cfg->n_reads == 0) return NSNP_OK;
    synth_fill_kernel<<<(unsigned)((cfg->n_reads + 63) / 64), 64, 0, (cudaStream_t)stream>>>(*cfg, cigar_off, seq_off, cigar, seq2, nmask);
    return nsnp::cuda_status("synth_fill_kernel");
}

}  // extern "C"
