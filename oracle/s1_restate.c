/*
 * ORACLE (test infrastructure, never on the product path).
 *
 * Plain-C restatement of NanoSNP's native s1 tools, from mpileup text to candidate windows:
 *   - TensorMaker::make_tensor          dna_sv_tensor/src/make_candidate_snp_tensor/tensor_maker.cpp:61-249
 *   - create_pileup_tensor              dna_sv_tensor/src/make_candidate_snp_tensor/main.cpp:113-312
 *   - s_load_next_vaf_info / make_predict_array   dna_sv_tensor/src/make_predict_data/main.cpp:76-127
 * Pinned against the reference's own binaries (oracle/_ref, compiled from /root/reference by
 * oracle/Makefile): tests/test_oracle_pinning.py byte-compares the .tensor and .pd files this code
 * writes with theirs on the SURVEY appendix C vectors and on seeded synthetic pileups.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_common.h"

#define NCH 18
#define MAX_INDEL 60                       /* tensor_maker.cpp:5 */

/* cpp_aux.cpp:85-102 nst_nt4_table: ACGTacgt -> 0..3, '-' -> 5, the rest 4 */
static int nt4(int c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
                 case 'T': case 't': return 3; case '-': return 5; default: return 4; }
}
/* tensor_maker.cpp:48-58 */
static int chan_of(int c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case '*': return 8;
                 case 'a': return 9; case 'c': return 10; case 'g': return 11; case 't': return 12; case '#': return 17;
                 default: return NCH; }
}
static int is_normal(int c) { return c && strchr("ACGTNacgtn*#", c) != NULL; }     /* tensor_maker.cpp:33 */
static int is_fwd(int c) { return c && strchr("ACGTN*", c) != NULL; }               /* tensor_maker.cpp:41 */

typedef struct { char sign; const char* seq; int len; int count; } covkey_t;       /* one cov_stats entry */
typedef struct { char* key; int count; } altent_t;                                  /* one alt_dict entry */

typedef struct {
    covkey_t* cov; int n_cov, cap_cov;
    altent_t* alt; int n_alt, cap_alt;
} scratch_t;

static void cov_add(scratch_t* s, char sign, const char* seq, int len) {
    for (int i = 0; i < s->n_cov; ++i) {
        covkey_t* k = &s->cov[i];
        if (k->sign == sign && k->len == len && memcmp(k->seq, seq, (size_t)len) == 0) { ++k->count; return; }
    }
    if (s->n_cov == s->cap_cov) { s->cap_cov = s->cap_cov ? 2 * s->cap_cov : 64; s->cov = realloc(s->cov, (size_t)s->cap_cov * sizeof *s->cov); }
    covkey_t k = {sign, seq, len, 1};
    s->cov[s->n_cov++] = k;
}
static void alt_add(scratch_t* s, const char* key, int count) {
    for (int i = 0; i < s->n_alt; ++i) if (strcmp(s->alt[i].key, key) == 0) { s->alt[i].count += count; return; }
    if (s->n_alt == s->cap_alt) { s->cap_alt = s->cap_alt ? 2 * s->cap_alt : 32; s->alt = realloc(s->alt, (size_t)s->cap_alt * sizeof *s->alt); }
    s->alt[s->n_alt].key = strdup(key); s->alt[s->n_alt].count = count; ++s->n_alt;
}
static void alt_clear(scratch_t* s) { for (int i = 0; i < s->n_alt; ++i) free(s->alt[i].key); s->n_alt = 0; }
static int alt_cmp(const void* a, const void* b) { return strcmp(((const altent_t*)a)->key, ((const altent_t*)b)->key); }

typedef struct { int depth; int pass_af; double af; } rowinfo_t;

/* make_tensor: column string -> 18 counts, depth, alt_dict (in s->alt, sorted), pass_af, af */
static rowinfo_t make_tensor(scratch_t* s, const uint8_t* ref, int64_t contig_len, int64_t off1, const char* b, int n,
                             double snp_min_af, double indel_min_af, int32_t* t)
{
    rowinfo_t ri;
    int rb = ref[off1 - 1];
    if (nt4(rb) >= 4) rb = isupper(rb) ? 'A' : 'a';                 /* evc_base_from, tensor_maker.hpp:38-44 */
    const int chr_base = toupper(rb);
    memset(t, 0, NCH * sizeof *t);
    s->n_cov = 0;
    /* tokenizer, tensor_maker.cpp:83-114 (b[n] == 0 like std::string::operator[](size())) */
    int i = 0;
    while (i < n) {
        const char c = b[i];
        if (c == '+' || c == '-') {
            ++i;
            int adv = 0;
            while (i <= n && isdigit((unsigned char)b[i])) { adv = adv * 10 + (b[i] - '0'); ++i; }
            if (adv <= MAX_INDEL) cov_add(s, c, b + i, adv);
            i += adv - 1;
        } else if (is_normal(c)) {
            cov_add(s, 0, b + i, 1);
        } else if (c == '^') {
            ++i;
        }
        ++i;
    }

    int max_ins0 = 0, max_del0 = 0, max_ins1 = 0, max_del1 = 0, depth = 0;
    alt_clear(s);
    /* pileup_dict keys in std::map order: A C D G I T */
    int pd[6] = {0, 0, 0, 0, 0, 0};
    static const char pd_key[6] = {'A', 'C', 'D', 'G', 'I', 'T'};
    char key[2 * MAX_INDEL + 8];
    for (int k = 0; k < s->n_cov; ++k) {
        const covkey_t* cv = &s->cov[k];
        const int count = cv->count;
        if (cv->sign == '+') {
            int m = 0; key[m++] = 'I'; key[m++] = (char)chr_base;
            for (int j = 0; j < cv->len; ++j) key[m++] = (char)toupper((unsigned char)cv->seq[j]);
            key[m] = 0; alt_add(s, key, count);
            pd[4] += count;
            if (is_fwd(cv->len ? cv->seq[0] : 0)) { t[NSNP_CH_I] += count; if (count > max_ins0) max_ins0 = count; }
            else { t[NSNP_CH_i] += count; if (count > max_ins1) max_ins1 = count; }
        } else if (cv->sign == '-') {
            int m = 0; key[m++] = 'D';
            for (int j = 1; j <= cv->len; ++j) {
                const int64_t q = off1 + j;                 /* 1-based; raw case (tensor_maker.cpp:151) */
                key[m++] = (q >= 1 && q <= contig_len) ? (char)ref[q - 1] : '?';
            }
            key[m] = 0; alt_add(s, key, count);
            pd[2] += count;
            if (is_fwd(cv->len ? cv->seq[0] : 0)) { t[NSNP_CH_D] += count; if (count > max_del0) max_del0 = count; }
            else { t[NSNP_CH_d] += count; if (count > max_del1) max_del1 = count; }
        } else {
            const int c0 = cv->seq[0];
            if (nt4(c0) < 4) {
                const int up = toupper(c0);
                pd[up == 'A' ? 0 : up == 'C' ? 1 : up == 'G' ? 3 : 5] += count;
                depth += count;
                if (up != chr_base) { key[0] = 'X'; key[1] = (char)up; key[2] = 0; alt_add(s, key, count); }
                t[chan_of(c0)] += count;
            } else if (c0 == '*') { t[NSNP_CH_STAR] += count; depth += count; }
            else if (c0 == '#') { t[NSNP_CH_POUND] += count; depth += count; }
        }
    }
    t[NSNP_CH_I1] = max_ins0; t[NSNP_CH_i1] = max_ins1; t[NSNP_CH_D1] = max_del0; t[NSNP_CH_d1] = max_del1;

    /* pileup_list: entries present in the map, stable-sorted by count descending (tensor_maker.cpp:195-199) */
    const int den = depth ? depth : 1;
    int ord[6], n_list = 0;
    for (int k = 0; k < 6; ++k) if (pd[k] > 0) ord[n_list++] = k;
    for (int a = 1; a < n_list; ++a) {                                /* insertion sort == libstdc++ for n <= 16 */
        const int v = ord[a]; int j = a - 1;
        while (j >= 0 && pd[v] > pd[ord[j]]) { ord[j + 1] = ord[j]; --j; }
        ord[j + 1] = v;
    }
    int pass_snp = 0, pass_indel = 0;
    int pass_af = n_list && pd_key[ord[0]] != chr_base;
    for (int a = 0; a < n_list; ++a) {
        const int k = ord[a], count = pd[k];
        if (pd_key[k] == chr_base) continue;
        if (pd_key[k] == 'I' || pd_key[k] == 'D') { pass_indel = pass_indel || (1.0 * count / den >= indel_min_af); continue; }
        pass_snp = pass_snp || (1.0 * count / den >= snp_min_af);
    }
    double af = n_list > 1 ? 1.0 * pd[ord[1]] / den : 0.0;
    if (n_list && pd_key[ord[0]] != chr_base) af = 1.0 * pd[ord[0]] / den;

    /* reference-channel overwrite, tensor_maker.cpp:230-246 */
    const int fsum = t[NSNP_CH_A] + t[NSNP_CH_C] + t[NSNP_CH_G] + t[NSNP_CH_T];
    t[chan_of(chr_base)] = -fsum;
    const int rsum = t[NSNP_CH_a] + t[NSNP_CH_c] + t[NSNP_CH_g] + t[NSNP_CH_t];
    t[chan_of(tolower(chr_base))] = -rsum;

    qsort(s->alt, (size_t)s->n_alt, sizeof *s->alt, alt_cmp);       /* std::map<string,int> iteration order */
    ri.depth = depth; ri.pass_af = pass_af || pass_snp || pass_indel; ri.af = af;
    return ri;
}

typedef struct { int32_t pos; int32_t depth; char* alt_info; } pending_t;

int64_t orc_s1_from_mpileup(const char* mpileup_path, const char* contig_name, const uint8_t* ref, int64_t contig_len,
                            double snp_min_af, double indel_min_af, int32_t min_coverage, int32_t flank,
                            orc_s1_out_t* out, const char* tensor_path, const char* pd_path)
{
    FILE* in = fopen(mpileup_path, "r");
    if (!in) return -1;
    FILE* ft = tensor_path ? fopen(tensor_path, "w") : NULL;
    FILE* fp = pd_path ? fopen(pd_path, "w") : NULL;
    if ((tensor_path && !ft) || (pd_path && !fp)) { fclose(in); return -1; }
    const int W = 2 * flank + 1;
    int32_t (*ring)[NCH] = malloc((size_t)W * sizeof *ring);
    pending_t* pend = NULL; int n_pend = 0, cap_pend = 0, head = 0;
    scratch_t sc; memset(&sc, 0, sizeof sc);
    char* line = NULL; size_t lcap = 0; ssize_t ll;
    int pos_offset = 0, num_filled = 0; int64_t pre = -1;
    int64_t n_out = 0;
    char* refsub = malloc((size_t)W + 1);
    char* tens = malloc((size_t)W * NCH * 14 + 16);

    while ((ll = getline(&line, &lcap, in)) >= 0) {
        while (ll > 0 && (line[ll - 1] == '\n' || line[ll - 1] == '\r')) line[--ll] = 0;
        /* split_line(line, "\t"): empty tokens are dropped (cpp_aux.cpp:44-59); columns 0,1,4 are used */
        char* col[8]; int clen[8]; int nc = 0;
        for (ssize_t i = 0; i < ll && nc < 8;) {
            while (i < ll && line[i] == '\t') ++i;
            if (i >= ll) break;
            const ssize_t st = i;
            while (i < ll && line[i] != '\t') ++i;
            col[nc] = line + st; clen[nc] = (int)(i - st); ++nc;
        }
        if (nc < 5) continue;
        col[1][clen[1]] = 0; col[4][clen[4]] = 0;
        const int64_t off1 = atoll(col[1]);
        if (off1 < 1 || off1 > contig_len) { n_out = -3; break; }                  /* hbn_assert main.cpp:170 */
        const int ref_base = toupper(ref[off1 - 1]);
        if (pre + 1 != off1) {                                                          /* main.cpp:174-178 */
            for (int k = head; k < n_pend; ++k) free(pend[k].alt_info);
            num_filled = 0; pos_offset = 0; head = n_pend = 0;
        }
        pre = off1;

        int32_t t[NCH];
        const rowinfo_t ri = make_tensor(&sc, ref, contig_len, off1, col[4], clen[4], snp_min_af, indel_min_af, t);
        if (out && out->counts) memcpy(out->counts + (off1 - 1) * NCH, t, sizeof t);
        const int cand = nt4(ref_base) < 4 && ri.pass_af && ri.depth >= min_coverage;      /* main.cpp:196 */
        if (out && out->flags) out->flags[off1 - 1] = (uint8_t)(1 | (cand ? 2 : 0));
        if (cand) {
            if (n_pend == cap_pend) { cap_pend = cap_pend ? 2 * cap_pend : 64; pend = realloc(pend, (size_t)cap_pend * sizeof *pend); }
            /* alt_info text: "<depth>-" then "KEY cnt " per alt_dict entry (main.cpp:225-231) */
            size_t need = 32; for (int k = 0; k < sc.n_alt; ++k) need += strlen(sc.alt[k].key) + 16;
            char* ai = malloc(need); int m = sprintf(ai, "%d-", ri.depth);
            for (int k = 0; k < sc.n_alt; ++k) m += sprintf(ai + m, "%s %d ", sc.alt[k].key, sc.alt[k].count);
            pend[n_pend].pos = (int32_t)off1; pend[n_pend].depth = ri.depth; pend[n_pend].alt_info = ai; ++n_pend;
        }
        memcpy(ring[pos_offset], t, sizeof t);
        ++num_filled;
        pos_offset = (pos_offset + 1) % W;
        if (n_pend > head && off1 - pend[head].pos == flank) {                         /* main.cpp:208 */
            pending_t c = pend[head++];
            if (num_filled >= W) {
                for (int k = 0; k < W; ++k) refsub[k] = (char)ref[c.pos - flank + k - 1];
                refsub[W] = 0;
                if (out && n_out < out->cand_cap) {
                    if (out->cand_pos) out->cand_pos[n_out] = c.pos;
                    if (out->cand_depth) out->cand_depth[n_out] = c.depth;
                    if (out->windows) {
                        int32_t* w = out->windows + n_out * W * NCH;
                        for (int k = 0; k < W; ++k) memcpy(w + k * NCH, ring[(pos_offset + k) % W], NCH * sizeof(int32_t));
                    }
                }
                if (ft || fp) {
                    int m = 0;
                    for (int k = 0; k < W; ++k) for (int j = 0; j < NCH; ++j) m += sprintf(tens + m, "%d ", ring[(pos_offset + k) % W][j]);
                    if (ft) fprintf(ft, "%s\t%d\t%s\t%s\t%s\n", contig_name, c.pos, refsub, tens, c.alt_info);
                    if (fp) {
                        /* make_predict_data/main.cpp:89-92,120-123: upper-case window, centre must be ACGT, trim alt_info */
                        char up[128]; for (int k = 0; k <= W; ++k) up[k] = (char)toupper((unsigned char)refsub[k]);
                        if (nt4(up[flank]) < 4) {
                            size_t al = strlen(c.alt_info); while (al && isspace((unsigned char)c.alt_info[al - 1])) --al;
                            fprintf(fp, "%s\t%s:%d:%s\t%.*s\n", tens, contig_name, c.pos, up, (int)al, c.alt_info);
                        }
                    }
                }
                ++n_out;
            }
            free(c.alt_info);
            if (head == n_pend) head = n_pend = 0;
        }
        if (n_pend == 0) head = 0;
    }
    for (int k = head; k < n_pend; ++k) free(pend[k].alt_info);
    alt_clear(&sc); free(sc.alt); free(sc.cov); free(pend); free(ring); free(line); free(refsub); free(tens);
    fclose(in); if (ft) fclose(ft); if (fp) fclose(fp);
    return n_out;
}
