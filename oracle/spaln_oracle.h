/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference's
 * DNA spliced-alignment DP kernels with quantised intron-length penalty
 * ("_wip" family, `spaln -A2/-A3`), at the canonical vector width of the
 * AVX2 build (nelem = 16 int16 lanes).
 *
 * Pinned against the unmodified reference compiled into oracle/_ref/ (see
 * tests/test_oracle_vs_reference.py and tests/golden/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this.
 *
 * Reference files restated (paths relative to /root/reference):
 *   src/fwd2s1_wip_simd.h:42-231   scoreonlyS1_wip
 *   src/fwd2s1_wip_simd.h:233-474  forwardS1_wip
 *   src/fwd2s1_simd.cc:163-262     fhinitS1 / fhlastS1
 *   src/fwd2s1_simd.h:179-182,191-333  checkpoint(), buffer layout
 *   src/rhomb_coord.h:65-235       Anti_rhomb_coord<CHAR> trace store + walk
 *   src/simd_functions.h:1010-1074 (AVX2 int16 ops), 61-151 (Vec* helpers)
 */
#ifndef SPALN_ORACLE_H
#define SPALN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_MAXQUANT 8

typedef struct {
    int32_t gop;        /* PwdB::BasicGOP (<0)            src/aln2.cc:106 */
    int32_t gep;        /* PwdB::BasicGEP (<0)            src/aln2.cc:107 */
    int32_t lgop;       /* PwdB::LongGOP                  src/aln2.cc:110 */
    int32_t lgep;       /* PwdB::LongGEP                  src/aln2.cc:108 */
    int32_t noll;       /* PwdB::Noll: 2 affine, 3 double affine */
    int32_t ipen;       /* IntronPenalty::Penalty() = GapWI  src/codepot.h:241 */
    int32_t llmt;       /* IntronPrm.llmt */
    int32_t nquant;     /* IntronPrm.nquant (1 under -A3) */
    int32_t quant_len[SO_MAXQUANT];   /* IntronPenalty::qm[j].len */
    int32_t quant_pen[SO_MAXQUANT];   /* IntronPenalty::qm[j].pen */
    int32_t avmch;      /* int(Simmtx::AvTrc())           src/fwd2s1_simd.h:198 */
    int32_t local;      /* algmode.lcl & 16 */
    int32_t spj;        /* b->inex.intr */
    int32_t simdim;     /* Simmtx::dim */
    const int32_t* simmtx;  /* dim x dim, row = query code, col = genome code */
    int32_t gappen1;    /* PwdB::GapPenalty(1) */
    /* inputs of the scalar exact-ILD kernel only (so_trcbk_ng); may be 0 / NULL otherwise */
    int32_t codonk1;        /* PwdB::codonk1 (GapExtPen: LongGEP beyond it) */
    int32_t n_penalty;      /* entries of penalty[] */
    const int16_t* penalty; /* IntronPenalty::Penalty(n), n in [0, n_penalty)  src/codepot.h:243-248 */
    const int16_t* sig53tab;/* Exinon::sig53tab[0][0..543]                      src/codepot.cc:281-285 */
} so_params;

typedef struct {
    const uint8_t* a;       /* query residue codes, a[i] == *a->at(i), i in [0, alen) */
    const uint8_t* b;       /* genome residue codes */
    const int16_t* sig5;    /* SGPT2::sig5 indexed by column n, n in [0, blen + 1] */
    const int16_t* sig3;    /* SGPT2::sig3 */
    int32_t a_left, a_right, b_left, b_right;   /* Seq::left / right */
    int32_t a_exgl, a_exgr, b_exgl, b_exgr;     /* INEX end-gap flags */
    int32_t lw, up;         /* WINDOW (width = up - lw + 3) */
    const uint16_t* int53;  /* scalar kernel only: INT53 nibbles by column n in [0, blen + 1]:
                               dinc5 | dinc3 << 4 | cano5 << 8 | cano3 << 12 (src/codepot.h:49-54) */
    const int32_t* cip;     /* scalar kernels only, may be NULL: Cip_score::cip_score(m) by query
                               position m in [0, a_right] (src/gsinfo.h:127-139; `sigB`,
                               src/fwd2s1.cc:254,338 and 1191,1262) */
} so_task;

/* forwardS1_wip: returns number of SKL corners written to skl[2*i], skl[2*i+1]
 * (m, n) in the order Anti_rhomb_coord::traceback emits them (end -> start),
 * or -1 on allocation failure.  *n_cells receives the inner-loop lane count. */
int so_forward_wip(const so_params* p, const so_task* t, int32_t* score,
                   int32_t* skl, int cap, int64_t* n_cells);

/* scoreonlyS1_wip */
int so_scoreonly_wip(const so_params* p, const so_task* t, int32_t* score);

/* hirschbergS1_wip (src/fwd2s1_wip_simd.h:476-864), single affine.  cpos holds
 * (n_im + 1) x 10 ints (Dim10 records); ranges receives a_left, a_right, b_left,
 * b_right as the reference leaves them in the Seq objects. */
int so_hirschberg_wip(const so_params* p, const so_task* t, int n_im,
                      int32_t* score, int32_t* cpos, int32_t* ranges);

/* Aln2s1::trcbkalignS_ng on its scalar branch (src/fwd2s1.cc:1667-1710): forwardS_ng
 * (217-444) with initS_ng / lastS_ng (141-215), Vmf::traceback (src/vmf.cc:125-140) and the
 * end-point adjustment; exact intron scoring SpJunc::spjscr (src/codepot.cc:74-77).  This is
 * what the reference runs for blocks with fewer than 8 query rows.  Returns the number of
 * corners, -1 allocation failure, -3 missing tables. */
int so_trcbk_ng(const so_params* p, const so_task* t, int32_t* score, int32_t* skl, int cap);

/* ---- splice-signal scan of a genomic DNA segment ---- */
typedef struct {
    int32_t rows, cols, offset, nalpha, morder;     /* PatMat, src/utilseq.h:62-90 */
    float tonic, min_elem;
    const float* mtx;                               /* rows x cols, column major (cols blocks of rows) */
} so_patmat;

typedef struct {
    so_patmat pat5, pat3;       /* EijPat::pattern5 / pattern3 */
    float fS, sss;              /* Exinon::fS, alprm2.sss */
    int32_t any;                /* algmode.any */
    const int16_t* sig53tab;    /* Exinon::sig53tab[0][0..543] */
} so_scan_params;

/* Exinon::intron53_c + intron53_n (src/codepot.cc:437-523) with PatMat::calcPatMat
 * (src/utilseq.cc:905-1002) for a sequence whose Exinon is built over [0, len):
 * codes[i] == *Seq::at(i); outputs indexed by column n in [0, len + 1]. */
void so_exinon_scan_n(const so_scan_params* sp, const uint8_t* codes, int len,
                      int16_t* sig5, int16_t* sig3, uint16_t* int53);

/* protein-side scan: Exinon::intron53_p (src/codepot.cc:525-619) over a TRON segment */
typedef struct {
    so_scan_params base;        /* pat5, pat3, fS, sss, any, sig53tab */
    so_patmat patI, patT;       /* EijPat::patternI (start codon), patternT (termination codon) */
    const float* codepot;       /* ExinPot::begin() of PwdB::codepot: [ndata][3], or NULL */
    int32_t ndata, cp_order;    /* 4^(order + 1), Markov order */
    float fact, z, bti, o;      /* Exinon::fact, alprm2.z, alprm2.bti, alprm2.o */
} so_scan_params_p;

/* out: 8 shorts per column n in [0, len + 1] (sig5, sig3, sigS, sigT, sigE, sigI, phs5, phs3) */
void so_exinon_scan_p(const so_scan_params_p* sp, const uint8_t* tron, int len, int16_t* out, uint16_t* int53);

/* Seq::nuc2tron (src/seq.cc:774-798): tron codes of a genomic DNA segment for protein queries;
 * codes points at at(0) and codes[-1], codes[len] must be readable (terminal residues) */
void so_nuc2tron(const uint8_t* gencode, const uint8_t* codes, int len, uint8_t* tron);

/* Aln2s1::hirschbergS_ng (src/fwd2s1.cc:764-1104) with hinitS_ng / hlastS_ng (701-762): the scalar
 * unidirectional Hirschberg pass of `-A0`.  imd_intvl: spacing of the intermediate rows (the member
 * lspS_ng sets, src/fwd2s1.cc:1839-1851); cpos: (n_im + 1) x 10 ints, entries [8], [9] = lowest /
 * highest diagonal of each block; ranges as for so_hirschberg_wip. */
int so_hirschberg_ng(const so_params* p, const so_task* t, int n_im, int imd_intvl,
                     int32_t* score, int32_t* cpos, int32_t* ranges);

/* Aln2s1::scorealoneS_ng (src/fwd2s1.cc:1163-1336): the scalar score-only kernel */
int so_scorealone_ng(const so_params* p, const so_task* t, int32_t* score);

/* Aln2s1::lspS_ng driver (trace-back vs multi-intermediate Hirschberg dispatch,
 * src/fwd2s1.cc:1801-1897) */
typedef struct {
    int32_t max_vmf_space;  /* MaxVmfSpace (-V), src/vmf.h:26 */
    int32_t sh;             /* alprm.sh band shoulder */
    int32_t ubh;            /* alprm.ubh: forced number of intermediates (0 = automatic) */
    int32_t alg;            /* algmode.alg (bit 2: recursive single-intermediate) */
} so_lsp_opts;

int so_lsp(const so_params* p, const so_task* t, const so_lsp_opts* o, int32_t* score,
           int32_t* skl, int cap, int* unsupported);

/* ---- protein x genome: SimdAln2h1::forwardH1_wip (src/fwd2h1_wip_simd.h:50-336) ---- */
typedef struct {
    int32_t gop, gep;       /* PwdB::BasicGOP, BasicGEP */
    int32_t lgep, codonk1;  /* PwdB::LongGEP, codonk1 (GapExtPen3) */
    int32_t gw1, gw2, gw3;  /* PwdB::GapW1, GapW2 (frame shifts), GapW3 (codon gap open + ext) */
    int32_t ipen, llmt, nquant;
    int32_t quant_len[SO_MAXQUANT], quant_pen[SO_MAXQUANT];
    int32_t avmch, local, lcl, spj, simdim;
    const int32_t* simmtx;  /* [aa code][tron code] */
    int32_t lgop;           /* PwdB::LongGOP (GapPenalty beyond codonk1; driver only) */
    int32_t gape1, gape2;   /* PwdB::GapE1, GapE2 (UnpPenalty3; driver only) */
    const struct so_ng_h_s* ng; /* driver only: inputs of the scalar kernel for blocks with fewer than 8
                                   rows (NULL: such blocks are reported as unsupported) */
} so_params_h;

typedef struct {
    const uint8_t* a;       /* amino-acid codes, a[i] == *a->at(i) */
    const uint8_t* b;       /* tron codes (Seq::nuc2tron, src/seq.cc:774-798), b[i] == *b->at(i) */
    const int16_t* sgpt6;   /* SGPT6 by column n in [0, b_len + 1]: sig5, sig3, sigS, sigT, sigE,
                               sigI, phs5, phs3 (src/codepot.h:34-43) */
    int32_t b_len;          /* Seq::len of the genomic segment (range of Exinon::good()) */
    int32_t a_left, a_right, b_left, b_right;
    int32_t a_exgl, a_exgr, b_exgl, b_exgr;     /* INEX flags, values 0..3 */
    int32_t lw, up;         /* WINDOW from stripe31 (width = up - lw + 7) */
    int32_t a_len;          /* Seq::len of the query (range check of mimd_postwork; driver only) */
    const int32_t* cip;     /* scalar kernel only, may be NULL: Cip_score::cip_score(c) by coding
                               position c = 3 m - phase in [0, 3 a_right + 1] (src/fwd2h1.cc:352-354) */
} so_task_h;

/* returns number of corners (want_trace) or 0; -1 allocation failure, -2 bad trace code */
int so_forward_h1_wip(const so_params_h* p, const so_task_h* t, int want_trace, int32_t* score,
                      int32_t* skl, int cap);

/* SimdAln2h1::hirschbergH1_wip (src/fwd2h1_wip_simd.h:338-773); outputs as so_hirschberg_wip */
int so_hirschberg_h1_wip(const so_params_h* p, const so_task_h* t, int n_im, int32_t* score,
                         int32_t* cpos, int32_t* ranges);

/* Aln2h1::lspH_ng driver (src/fwd2h1.cc:2134-2230) with trcbkalignH_ng (SIMD branch),
 * diagonalH_ng, mimd_postwork, rcsv_postwork and stripe31 */
int so_lsp_h(const so_params_h* p, const so_task_h* t, const so_lsp_opts* o, int32_t* score,
             int32_t* skl, int cap, int* unsupported);

/* extra inputs of the scalar protein kernel */
typedef struct so_ng_h_s {
    const int16_t* penalty; int32_t n_penalty;      /* IntronPenalty::Penalty(len) */
    const int16_t* sig53tab;                        /* Exinon::sig53tab[0][0..543] */
    const uint16_t* int53;                          /* INT53 nibbles by column */
    const uint8_t* spj_tabs;    /* spj_tron_tab[257][2] | spj_amb_tron_tab[64][2] | spj_tron_amb_tab[64][2] | aa2nuc[26] */
    int32_t minl, extragop, gw3l, noll;             /* IntronPrm.minl, PwdB::ExtraGOP, GapW3L, Noll */
} so_ng_h;

/* Aln2h1::trcbkalignH_ng on its scalar branch (src/fwd2h1.cc:1997-2041): forwardH_ng (294-617)
 * with initH_ng / lastH_ng (143-292), Vmf trace-back and the end-point adjustment */
int so_trcbk_h_ng(const so_params_h* p, const so_ng_h* x, const so_task_h* t, int32_t* score,
                  int32_t* skl, int cap);

/* Aln2h1::hirschbergH_ng (src/fwd2h1.cc:1085-1520) with hinitH_ng / hlastH_ng (941-1083): the scalar
 * protein Hirschberg pass of `-A0`.  imd_intvl: spacing of the intermediate rows (lspH_ng,
 * src/fwd2h1.cc:2170-2183); cpos: (n_im + 1) x 10 ints, [8], [9] = diagonal bounds of each block. */
int so_hirschberg_h_ng(const so_params_h* p, const struct so_ng_h_s* x, const so_task_h* t, int n_im,
                       int imd_intvl, int32_t* score, int32_t* cpos, int32_t* ranges);

#ifdef __cplusplus
}
#endif
#endif
