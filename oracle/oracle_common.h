/* ORACLE (test infrastructure).  Shared declarations of the C restatement. */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H
#include <stdint.h>
#include "../include/nanosnp_b200.h"   /* only for the flat-array input type nsnp_reads_t / channel enum */

int64_t orc_mpileup_write(const nsnp_reads_t* r, const char* contig_name, int32_t min_mapq, uint32_t excl_flags,
                          int32_t max_depth, const char* path, int append, int32_t* max_depth_seen);

int64_t orc_depth_cap(const nsnp_reads_t* r, int32_t min_mapq, uint32_t excl_flags, int32_t max_depth, uint8_t* dropped);

/* s1 restatement: mpileup text -> counts / flags / candidates / windows / .tensor text */
typedef struct orc_s1_out {
    int32_t* counts;      /* optional [contig_len][18]: the 18-channel row of every mpileup row (else 0) */
    uint8_t* flags;       /* optional [contig_len]: bit0 row exists, bit1 candidate gate passed */
    int32_t* cand_pos;    /* optional [cand_cap]: 1-based centre positions in emit order */
    int32_t* windows;     /* optional [cand_cap][33][18] */
    int32_t* cand_depth;  /* optional [cand_cap] */
    int64_t  cand_cap;
} orc_s1_out_t;

int64_t orc_s1_from_mpileup(const char* mpileup_path, const char* contig_name, const uint8_t* ref, int64_t contig_len,
                            double snp_min_af, double indel_min_af, int32_t min_coverage, int32_t flank,
                            orc_s1_out_t* out, const char* tensor_path, const char* pd_path);
#endif
