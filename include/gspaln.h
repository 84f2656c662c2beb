/* gspaln.h -- C-ABI of the B200-native spliced-alignment DP engine.
 *
 * Drop-in boundary for the DP hot path of ogotoh/spaln ("the reference").
 * Every entry point is `extern "C"`, takes plain pointers and sizes, returns
 * 0 on success or a negative GSPALN_E* code; nothing throws, no globals are
 * read after gspaln_create().  All paths are relative to /root/reference.
 *
 * What each entry point replaces in the reference:
 *
 *   gspaln_create            the process-global parameter set the kernels read
 *                            (PwdB src/aln.h:235-308 built at src/aln2.cc:99-137,
 *                            IntronPrm src/codepot.cc:38-46 + IntronPenalty::qm
 *                            src/codepot.h:218-257, Simmtx::mtx src/simmtx.h:37-66,
 *                            algmode.lcl src/clib.h:38-56), frozen into one POD.
 *   gspaln_submit            kind GSPALN_FORWARD_WIP  = SimdAln2s1 ctor +
 *                            SimdAln2s1::forwardS1_wip(Mfile*)
 *                            (src/fwd2s1_simd.h:191-333, src/fwd2s1_wip_simd.h:233-474)
 *                            as called from Aln2s1::trcbkalignS_ng
 *                            (src/fwd2s1.cc:1676-1689, mode == 1);
 *                            kind GSPALN_SCOREONLY_WIP = SimdAln2s1::scoreonlyS1_wip()
 *                            (src/fwd2s1_wip_simd.h:42-231) as called from
 *                            HomScoreS_ng (src/fwd2s1.cc:2696-2716).
 *                            One gspaln_task == one such call; a batch of tasks
 *                            is what the reference's pthread workers
 *                            (src/spaln.cc:1363-1468) issue one at a time.
 *   gspaln_result.skl        the (m, n) corner records the reference appends to
 *                            the caller's Mfile through
 *                            Anti_rhomb_coord::traceback (src/rhomb_coord.h:222-235),
 *                            same order (alignment end first).
 *
 * Arithmetic is the reference's int16 saturating lane arithmetic at the
 * canonical AVX2 width (strips of 16 query rows); results are bit-identical.
 */
#ifndef GSPALN_H
#define GSPALN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSPALN_MAXQUANT 8
#define GSPALN_MAXDIM   32      /* Simmtx::dim upper bound handled (DNA: 17) */

enum {
    GSPALN_OK = 0,
    GSPALN_EINVAL = -1,         /* bad argument / unsupported parameter */
    GSPALN_ENOMEM = -2,         /* device or host allocation failed */
    GSPALN_ECUDA = -3,          /* CUDA runtime error (see gspaln_last_error) */
    GSPALN_ENODEV = -4          /* no usable CUDA device: there is NO CPU fallback */
};

enum {                          /* gspaln_task.kind */
    GSPALN_FORWARD_WIP = 0,     /* score + trace-back corners */
    GSPALN_SCOREONLY_WIP = 1,   /* score only */
    GSPALN_HIRSCHBERG_WIP = 2,  /* SimdAln2s1::hirschbergS1_wip(Dim10* cpos, n_imd)
                                   (src/fwd2s1_wip_simd.h:476-864): score, crossing records
                                   and the narrowed sequence ranges; global / semi-global only */
    GSPALN_FORWARD_NG = 3,      /* Aln2s1::trcbkalignS_ng on its scalar branch: forwardS_ng
                                   (src/fwd2s1.cc:217-444) + Vmf::traceback + end adjustment
                                   (1667-1710), exact intron scoring.  The reference runs it for
                                   blocks with fewer than 8 query rows.  Needs
                                   gspaln_set_ng_tables() and gspaln_task.int53. */
    GSPALN_SCOREALONE_NG = 4,   /* Aln2s1::scorealoneS_ng (src/fwd2s1.cc:1163-1336): the scalar
                                   score-only kernel HomScoreS_ng runs under -A0 and for queries
                                   shorter than 4 residues (src/fwd2s1.cc:2704-2705).  Same tables. */
    GSPALN_HIRSCHBERG_NG = 5    /* Aln2s1::hirschbergS_ng (src/fwd2s1.cc:764-1104): the scalar Hirschberg
                                   pass of -A0 with exact intron scoring.  n_imd = the number of
                                   intermediate rows lspS_ng asks for BEFORE its even-division
                                   correction (src/fwd2s1.cc:1850-1851; spacing (m + n_imd) / (n_imd + 1),
                                   one row fewer when it divides the query evenly).  Results as for
                                   GSPALN_HIRSCHBERG_WIP; cpos[i][8], [9] = lowest / highest diagonal of
                                   block i (the band of its re-alignment).  Same tables as FORWARD_NG. */
};

enum {                          /* gspaln_result.status */
    GSPALN_ST_OK = 0,
    GSPALN_ST_SKL_OVERFLOW = 1, /* more corners than skl_cap: n_skl is the needed count */
    GSPALN_ST_BAD_TRACE = 2,    /* reference would have called fatal("Unexpected dir") */
    GSPALN_ST_UNSUPPORTED = 3,  /* needs a kernel that is not on the device yet */
    GSPALN_ST_INTERNAL = 4,     /* device-side scheduling guard tripped (a bug: please report) */
    GSPALN_ST_VMF_OVERFLOW = 5  /* GSPALN_FORWARD_NG: more path records than the workspace holds
                                   (the reference's fatal("Too many Vmf records")) */
};

/* frozen scoring parameters (reference globals -> one POD) */
typedef struct gspaln_params {
    int32_t gop;                /* PwdB::BasicGOP  (< 0) */
    int32_t gep;                /* PwdB::BasicGEP  (< 0) */
    int32_t lgop;               /* PwdB::LongGOP */
    int32_t lgep;               /* PwdB::LongGEP */
    int32_t noll;               /* PwdB::Noll: 2 = affine, 3 = double affine (forward / score-only kernels) */
    int32_t ipen;               /* IntronPenalty::Penalty() == GapWI */
    int32_t llmt;               /* IntronPrm.llmt */
    int32_t nquant;             /* IntronPrm.nquant (1 for -A3) */
    int32_t quant_len[GSPALN_MAXQUANT];     /* IntronPenalty::qm[j].len */
    int32_t quant_pen[GSPALN_MAXQUANT];     /* IntronPenalty::qm[j].pen */
    int32_t avmch;              /* int(Simmtx::AvTrc()) */
    int32_t local;              /* algmode.lcl & 16 */
    int32_t spj;                /* Seq::inex.intr of the genomic sequence */
    int32_t simdim;             /* Simmtx::dim */
    int32_t gappen1;            /* PwdB::GapPenalty(1) */
    int32_t simmtx[GSPALN_MAXDIM * GSPALN_MAXDIM];  /* mtx[q][g] at [q * simdim + g] */
} gspaln_params;

/* one DP problem == one SimdAln2s1 construction + kernel call */
typedef struct gspaln_task {
    int32_t kind;
    const uint8_t* a;           /* query codes; a[i] == *Seq::at(i), i in [a_left, a_right) */
    const uint8_t* b;           /* genome codes; b[i] == *Seq::at(i), i in [b_left, b_right) */
    const int16_t* sig5;        /* Exinon::data_n[n].sig5 by column n, n in [b_left, b_right] */
    const int16_t* sig3;        /* Exinon::data_n[n].sig3 */
    int32_t a_left, a_right;    /* Seq::left / Seq::right of the query */
    int32_t b_left, b_right;    /* ... of the genomic segment */
    int32_t a_exgl, a_exgr;     /* INEX::exgl / exgr (src/seq.h:148-172) */
    int32_t b_exgl, b_exgr;
    int32_t lw, up;             /* WINDOW (src/cmn.h:133); width = up - lw + 3 */
    int32_t skl_cap;            /* capacity of result.skl in corners */
    int32_t n_imd;              /* GSPALN_HIRSCHBERG_WIP / _NG: number of intermediate rows (>= 1) */
    const uint16_t* int53;      /* GSPALN_FORWARD_NG (and gspaln_lsp blocks with < 8 rows), else may be
                                   NULL: Exinon::int53[n] by column n (src/codepot.h:49-54) as
                                   dinc5 | dinc3 << 4 | cano5 << 8 | cano3 << 12 */
    const int32_t* cip;         /* optional (NULL: none): Cip_score::cip_score(m) by query position m in
                                   [0, a_right] (src/gsinfo.h:127-139), the bonus for intron positions
                                   conserved with the query's own annotation.  Added at acceptors by the
                                   exact-ILD kernels (`sigB`, src/fwd2s1.cc:254,338 and 1191,1262); the
                                   `_wip` kernels of the reference do not read it, nor do ours */
} gspaln_task;

typedef struct gspaln_result {
    int32_t score;              /* VTYPE return value of the kernel */
    int32_t status;
    int32_t n_skl;              /* corners written (or needed on overflow) */
    int32_t reserved;
    int64_t cells;              /* query rows x band columns evaluated (throughput accounting) */
    int32_t* skl;               /* caller buffer: skl[2*i] = m, skl[2*i+1] = n */
    int32_t ranges[4];          /* HIRSCHBERG: a.left, a.right, b.left, b.right as the reference
                                   leaves them in the Seq objects (src/fwd2s1_wip_simd.h:812-861) */
    int32_t* cpos;              /* HIRSCHBERG: caller buffer, (n_imd + 1) x 10 int32 (Dim10 records,
                                   src/udh_intermediate.h:90; rows end with INT_MAX - 2) */
} gspaln_result;

typedef struct gspaln_ctx gspaln_ctx;

/* timing of the last submit, CUDA events on the engine's own stream */
typedef struct gspaln_timing {
    float h2d_ms;               /* host->device copies */
    float kernel_ms;            /* DP kernel(s), start to end */
    float d2h_ms;               /* device->host copies */
    int32_t launches;           /* kernels launched by this call */
    int64_t h2d_bytes, d2h_bytes;
    int64_t trace_bytes;        /* trace-code bytes the DP kernel wrote */
    int64_t cells;
} gspaln_timing;

int  gspaln_create(gspaln_ctx** out, const gspaln_params* prm, int device);
void gspaln_destroy(gspaln_ctx* ctx);

/* Tables of the exact intron scoring (SpJunc::spjscr, src/codepot.cc:74-77) used by
 * GSPALN_FORWARD_NG: sig53tab = Exinon::sig53tab[0][0 .. 543] (src/codepot.cc:281-285),
 * penalty[len] = IntronPenalty::Penalty(len) for len in [0, n_penalty) (src/codepot.h:243-248;
 * evaluated by the caller so that the float tail of the distribution is the reference's own
 * libm result), codonk1 = PwdB::codonk1 (GapExtPen).  alprm2.Z must be 0 (no intron potential).
 * Problems whose genomic range is n_penalty or longer are refused. */
int  gspaln_set_ng_tables(gspaln_ctx* ctx, const int16_t* sig53tab, const int16_t* penalty,
                          int32_t n_penalty, int32_t codonk1);

/* blocking: pack + H2D + kernels + D2H.  results[i].skl must point to
 * tasks[i].skl_cap * 2 int32 (may be NULL for score-only tasks). */
int  gspaln_submit(gspaln_ctx* ctx, const gspaln_task* tasks, int n,
                   gspaln_result* results);

/* split form: inputs resident in HBM before the timed region starts.
 * gspaln_upload packs and copies the batch; gspaln_run launches the kernels
 * (may be called repeatedly on the same resident batch); gspaln_download
 * fetches scores and corners. */
int  gspaln_upload(gspaln_ctx* ctx, const gspaln_task* tasks, int n);
int  gspaln_run(gspaln_ctx* ctx);
int  gspaln_download(gspaln_ctx* ctx, gspaln_result* results);

/* The DP driver: Aln2s1::lspS_ng (src/fwd2s1.cc:1801-1897) over a batch.  Each task is one
 * lspS_ng call (task.kind is ignored).  The driver decides per problem between trace-back
 * (trcbkalignS_ng, src/fwd2s1.cc:1667-1710) and the multi-intermediate unidirectional
 * Hirschberg method with the reference's own space estimate, runs the Hirschberg passes and
 * the block re-alignments of mimd_postwork / rcsv_postwork (src/fwd2s1.cc:1714-1799) as
 * further device batches, and returns the corner list in the order the reference writes
 * it to its Mfile.  Blocks with fewer than 8 query rows go to the scalar kernel
 * (GSPALN_FORWARD_NG) like in the reference; without gspaln_set_ng_tables() / task.int53 they
 * set GSPALN_ST_UNSUPPORTED on that problem, as does the double-affine Hirschberg route. */
typedef struct gspaln_lsp_opts {
    int32_t max_vmf_space;      /* MaxVmfSpace (-V; default 32 MiB, src/vmf.h:26) */
    int32_t sh;                 /* alprm.sh: band shoulder for the block re-alignments */
    int32_t ubh;                /* alprm.ubh: forced number of intermediates (0 = automatic) */
    int32_t alg;                /* algmode.alg (bit 2: recursive single-intermediate method) */
} gspaln_lsp_opts;

int  gspaln_lsp(gspaln_ctx* ctx, const gspaln_task* tasks, int n, const gspaln_lsp_opts* opts,
                gspaln_result* results);

/* Coalescing queue for the literal drop-in: Spaln issues its DP problems one at a time from each
 * pthread worker (src/spaln.cc:1363-1468).  Workers call gspaln_queue_submit() with ONE task and
 * block until its result is there; a dispatcher thread owned by the queue collects what the
 * workers have queued (up to max_batch tasks, waiting at most max_wait_us for stragglers once
 * the first task has arrived) and runs it as one gspaln_submit() batch.  Thread-safe; the
 * context must not be used directly while a queue is attached to it. */
typedef struct gspaln_queue gspaln_queue;
int  gspaln_queue_create(gspaln_queue** out, gspaln_ctx* ctx, int max_batch, int max_wait_us);
int  gspaln_queue_submit(gspaln_queue* q, const gspaln_task* task, gspaln_result* result);
/* the same for whole driver calls: one Aln2s1::lspS_ng call (src/fwd2s1.cc:1801-1897) per worker,
 * coalesced into gspaln_lsp() batches (calls with equal options share a batch) */
int  gspaln_queue_submit_lsp(gspaln_queue* q, const gspaln_task* task, const gspaln_lsp_opts* opts,
                             gspaln_result* result);
int  gspaln_queue_stats(const gspaln_queue* q, int64_t* tasks, int64_t* batches);
void gspaln_queue_destroy(gspaln_queue* q);

int  gspaln_get_timing(const gspaln_ctx* ctx, gspaln_timing* out);
const char* gspaln_last_error(const gspaln_ctx* ctx);
int  gspaln_device_count(void);
const char* gspaln_version(void);

/* band cells of a task exactly as the scalar reference counts them
 * (inner-loop trip count of Aln2s1::forwardS_ng, src/fwd2s1.cc:252-276) */
int64_t gspaln_task_cells(const gspaln_task* t);

/* ======================================================================================
 * Splice-signal scan of a genomic DNA segment: what the Exinon
 * constructor computes per (segment, strand) before any DP runs --
 *   Exinon::intron53_c  (src/codepot.cc:437-477)  INT53: dinucleotide codes + site classes
 *   Exinon::intron53_n  (src/codepot.cc:479-523)  SGPT2 sig5 / sig3 from the two splice PSSMs
 *   PatMat::calcPatMat  (src/utilseq.cc:905-1002) Markov order <= 2, Seq::many == 1
 * One thread per genome position; same fp32 operation order as the reference, so the shorts
 * are bit-identical.  Outputs are indexed by column n in [0, len + 1] like Exinon::data_n
 * built over [0, len); sig3[0] and sig5[len - 1] are undefined in the reference
 * (uninitialised INT53 entries) and use dinucleotide code 0 here.
 * ====================================================================================== */
typedef struct gspaln_patmat {      /* PatMat, src/utilseq.h:62-90 */
    int32_t rows, cols, offset, nalpha, morder;
    float tonic, min_elem;
    const float* mtx;               /* cols blocks of rows floats */
} gspaln_patmat;

typedef struct gspaln_scan_params {
    gspaln_patmat pat5, pat3;       /* EijPat::pattern5 / pattern3 (mtx may be NULL: no PSSM) */
    float fS, sss;                  /* Exinon::fS, alprm2.sss */
    int32_t any;                    /* algmode.any */
    int16_t sig53tab[32];           /* Exinon::sig53tab[0][0..15] (5') and [1][0..15] (3') */
} gspaln_scan_params;

typedef struct gspaln_scan gspaln_scan;

int  gspaln_scan_create(gspaln_scan** out, const gspaln_scan_params* prm, int device);
void gspaln_scan_destroy(gspaln_scan* sc);
/* blocking, host buffers: codes[i] == *Seq::at(i), i in [0, len); outputs hold len + 2 entries */
int  gspaln_exinon_scan(gspaln_scan* sc, const uint8_t* codes, int64_t len,
                        int16_t* sig5, int16_t* sig3, uint16_t* int53);
/* split form (segment resident in HBM; gspaln_scan_run may be repeated) */
int  gspaln_scan_upload(gspaln_scan* sc, const uint8_t* codes, int64_t len);
int  gspaln_scan_run(gspaln_scan* sc);
int  gspaln_scan_download(gspaln_scan* sc, int16_t* sig5, int16_t* sig3, uint16_t* int53);
/* ms of the last upload / run / download (CUDA events on the engine's stream) */
int  gspaln_scan_get_timing(const gspaln_scan* sc, float* h2d_ms, float* kernel_ms, float* d2h_ms);
const char* gspaln_scan_last_error(const gspaln_scan* sc);

/* Protein-side scan: Exinon::intron53_p (src/codepot.cc:525-619) over a TRON segment (the output
 * of gspaln_nuc2tron): per column the SGPT6 record -- sig5 / sig3 as above, sigS / sigT from the
 * start- and termination-codon PSSMs (EijPat::patternI / patternT), sigE from the coding potential
 * (ExinPot::calcScr_3, src/utilseq.cc:1423-1459: 5th-order Markov model in three phases) with the
 * termination-codon adjustments, and the intron phases phs5 / phs3 -- plus INT53.  A tron code
 * keeps the middle nucleotide of its codon (tnredctab, src/seq.cc:41), which is what the PSSMs and
 * the potential read.  Limits of this version: algmode.any == 0, no branch-point PSSM, no intron
 * potential (alprm2.Z == 0), Seq::many == 1; entries that depend on INT53 halves the reference
 * leaves uninitialised (sig3 / phs3 of columns 0-1, sig5 / phs5 of the last two columns) use
 * dinucleotide code 0. */
typedef struct gspaln_scan_params_p {
    gspaln_scan_params base;
    gspaln_patmat patI, patT;       /* mtx may be NULL */
    const float* codepot;           /* ExinPot::begin() of PwdB::codepot, [ndata][3]; may be NULL */
    int32_t ndata, cp_order;        /* 4^(order + 1), Markov order */
    float fact, z, bti, o;          /* Exinon::fact, alprm2.z, alprm2.bti, alprm2.o */
} gspaln_scan_params_p;

int  gspaln_scan_create_p(gspaln_scan** out, const gspaln_scan_params_p* prm, int device);
/* blocking, host buffers: tron[i] == *Seq::at(i) of the TRON segment; sg and int53 hold len + 2 entries */
int  gspaln_exinon_scan_p(gspaln_scan* sc, const uint8_t* tron, int64_t len,
                          struct gspaln_sgpt6* sg, uint16_t* int53);

/* Seq::nuc2tron (src/seq.cc:774-798, nuc2tron3 src/utilseq.cc:205-224; Seq::many == 1): the "tron"
 * residues a protein query is aligned against -- position i becomes the translation of the codon
 * (i - 1, i, i + 1) with the genetic code table `gencode` (src/utilseq.cc:38).  codes holds
 * at(-1 .. len), i.e. len + 2 bytes with the two terminal residues; tron receives at(0 .. len - 1).
 * Blocking, host buffers; *kernel_ms (may be NULL) receives the CUDA-event time of one kernel run
 * with the segment resident.  There is no CPU fallback (GSPALN_ENODEV without a device). */
int  gspaln_nuc2tron(int device, const uint8_t* gencode, const uint8_t* codes, int64_t len,
                     uint8_t* tron, float* kernel_ms);

/* ======================================================================================
 * Protein query x genomic segment: SimdAln2h1 (src/fwd2h1_simd.h:69-382).
 *
 *   gspaln_h_create      freezes what SimdAln2h1::forwardH1_wip / fhinitH1 / fhlastH1 read from
 *                        PwdB (GapW1/W2/W3, BasicGOP/GEP, LongGEP, codonk1: src/aln.h:235-308),
 *                        IntronPrm + IntronPenalty::qm, Simmtx::mtx[aa][tron], algmode.lcl.
 *   gspaln_h_submit      kind GSPALN_FORWARD_WIP   = SimdAln2h1 ctor + forwardH1_wip(Mfile*)
 *                        (src/fwd2h1_wip_simd.h:50-336) as called from Aln2h1::trcbkalignH_ng
 *                        (src/fwd2h1.cc:2006-2019); kind GSPALN_SCOREONLY_WIP = forwardH1_wip(0)
 *                        as called from HomScoreH_ng (src/fwd2h1.cc:3293-3307).
 *   gspaln_result.skl    the corners Anti_rhomb_coord<SHORT>::traceback (step 3,
 *                        src/rhomb_coord.h:222-235) appends to the caller's Mfile.
 * ====================================================================================== */

/* SGPT6, src/codepot.h:34-43 (same layout: 6 shorts + 2 chars, 14 bytes) */
typedef struct gspaln_sgpt6 {
    int16_t sig5, sig3, sigS, sigT, sigE, sigI;
    int8_t phs5, phs3;
} gspaln_sgpt6;

typedef struct gspaln_h_params {
    int32_t gop;                /* PwdB::BasicGOP */
    int32_t gep;                /* PwdB::BasicGEP */
    int32_t lgep;               /* PwdB::LongGEP  (GapExtPen3 beyond codonk1) */
    int32_t codonk1;            /* PwdB::codonk1 */
    int32_t gw1, gw2, gw3;      /* PwdB::GapW1, GapW2 (frame shifts), GapW3 (codon gap) */
    int32_t ipen;               /* IntronPenalty::Penalty() == GapWI */
    int32_t llmt;               /* IntronPrm.llmt */
    int32_t nquant;             /* IntronPrm.nquant */
    int32_t quant_len[GSPALN_MAXQUANT];
    int32_t quant_pen[GSPALN_MAXQUANT];
    int32_t avmch;              /* int(Simmtx::AvTrc()) */
    int32_t lcl;                /* algmode.lcl (bit 4: local, bit 1: termination-codon bonus) */
    int32_t spj;                /* Seq::inex.intr of the genomic sequence */
    int32_t simdim;             /* row stride of simmtx below */
    int32_t simmtx[GSPALN_MAXDIM * GSPALN_MAXDIM];  /* mtx[aa][tron] at [aa * simdim + tron] */
    int32_t lgop;               /* PwdB::LongGOP          (driver: GapPenalty of an all-gap problem) */
    int32_t gape1, gape2;       /* PwdB::GapE1, GapE2     (driver: UnpPenalty3) */
} gspaln_h_params;

typedef struct gspaln_h_task {
    int32_t kind;               /* GSPALN_FORWARD_WIP, GSPALN_SCOREONLY_WIP, GSPALN_HIRSCHBERG_WIP
                                   (SimdAln2h1::hirschbergH1_wip, src/fwd2h1_wip_simd.h:338-773),
                                   GSPALN_FORWARD_NG (scalar forwardH_ng, see gspaln_h_set_ng_tables) or
                                   GSPALN_HIRSCHBERG_NG (scalar Aln2h1::hirschbergH_ng, src/fwd2h1.cc:1085-1520:
                                   the Hirschberg pass of -A0; n_imd as for the DNA kind) */
    const uint8_t* a;           /* amino-acid codes; a[i] == *Seq::at(i) */
    const uint8_t* b;           /* tron codes (Seq::nuc2tron, src/seq.cc:774-798); b[i] == *Seq::at(i) */
    const gspaln_sgpt6* sg;     /* Exinon::data_p[n], n in [0, b_len + 1] */
    int32_t b_len;              /* Seq::len of the genomic segment (range of Exinon::good()) */
    int32_t a_left, a_right, b_left, b_right;
    int32_t a_exgl, a_exgr, b_exgl, b_exgr;     /* INEX values 0..3 */
    int32_t lw, up;             /* WINDOW from stripe31 (src/aln2.cc:178-199); width = up - lw + 7 */
    int32_t skl_cap;
    int32_t n_imd;              /* GSPALN_HIRSCHBERG_WIP / _NG: number of intermediate rows (>= 1) */
    int32_t a_len;              /* Seq::len of the query (driver: range check of mimd_postwork) */
    const uint16_t* int53;      /* GSPALN_FORWARD_NG (and gspaln_h_lsp blocks with < 8 rows), else may be
                                   NULL: Exinon::int53[n] by column, as in gspaln_task.int53 */
    const int32_t* cip;         /* optional (NULL: none): Cip_score::cip_score(c) by coding position
                                   c = 3 m - phase in [0, 3 a_right + 1] (src/gsinfo.h:127-139), added
                                   at acceptors by the exact-ILD kernel (sigB[phs], src/fwd2h1.cc:352-354,
                                   483); the `_wip` kernels do not read it */
} gspaln_h_task;

typedef struct gspaln_h_ctx gspaln_h_ctx;

int  gspaln_h_create(gspaln_h_ctx** out, const gspaln_h_params* prm, int device);
void gspaln_h_destroy(gspaln_h_ctx* ctx);
int  gspaln_h_submit(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n, gspaln_result* results);

/* Tables of the scalar kernel GSPALN_FORWARD_NG for protein queries: Aln2h1::trcbkalignH_ng on its
 * scalar branch (src/fwd2h1.cc:1997-2041): forwardH_ng (294-617) + Vmf::traceback + end adjustment,
 * what the reference runs for blocks with fewer than 8 query rows.  sig53tab, penalty as in
 * gspaln_set_ng_tables; spj_tabs = spj_tron_tab[257][2] | spj_amb_tron_tab[64][2] |
 * spj_tron_amb_tab[64][2] (src/codepot.h:130-190) | aa2nuc[26] (src/seq.cc:76), 796 bytes;
 * minl = IntronPrm.minl, extragop = PwdB::ExtraGOP, gw3l = PwdB::GapW3L, noll = PwdB::Noll. */
int  gspaln_h_set_ng_tables(gspaln_h_ctx* ctx, const int16_t* sig53tab, const int16_t* penalty,
                            int32_t n_penalty, const uint8_t* spj_tabs, int32_t minl,
                            int32_t extragop, int32_t gw3l, int32_t noll);
int  gspaln_h_upload(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n);
int  gspaln_h_run(gspaln_h_ctx* ctx);
int  gspaln_h_download(gspaln_h_ctx* ctx, gspaln_result* results);
int  gspaln_h_get_timing(const gspaln_h_ctx* ctx, gspaln_timing* out);
const char* gspaln_h_last_error(const gspaln_h_ctx* ctx);
/* The protein driver: Aln2h1::lspH_ng (src/fwd2h1.cc:2134-2230) over a batch; each task is one
 * lspH_ng call (task.kind is ignored).  Same scheme as gspaln_lsp: trivial problems and the
 * single-diagonal case (diagonalH_ng, 1963-1995) on the host, trace-back problems
 * (trcbkalignH_ng, 1997-2041, SIMD branch), Hirschberg passes (hirschbergH1_wip) and the block
 * re-alignments of mimd_postwork / rcsv_postwork (2045-2132, re-banded with stripe31) as device
 * batches, one per recursion level.  Blocks with fewer than 8 query rows go to the scalar kernel
 * (GSPALN_FORWARD_NG: forwardH_ng) like in the reference; without gspaln_h_set_ng_tables() /
 * task.int53 they set GSPALN_ST_UNSUPPORTED. */
int  gspaln_h_lsp(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n, const gspaln_lsp_opts* opts,
                  gspaln_result* results);
/* coalescing queue of the protein path: as gspaln_queue_* (Aln2h1::trcbkalignH_ng / lspH_ng calls
 * of the pthread workers, src/spaln.cc:1363-1468) */
typedef struct gspaln_h_queue gspaln_h_queue;
int  gspaln_h_queue_create(gspaln_h_queue** out, gspaln_h_ctx* ctx, int max_batch, int max_wait_us);
int  gspaln_h_queue_submit(gspaln_h_queue* q, const gspaln_h_task* task, gspaln_result* result);
int  gspaln_h_queue_submit_lsp(gspaln_h_queue* q, const gspaln_h_task* task, const gspaln_lsp_opts* opts,
                               gspaln_result* result);
int  gspaln_h_queue_stats(const gspaln_h_queue* q, int64_t* tasks, int64_t* batches);
void gspaln_h_queue_destroy(gspaln_h_queue* q);
/* amino acid x nucleotide band cells as the scalar reference counts them
 * (Aln2h1::forwardH_ng inner loop, src/fwd2h1.cc:326-331) */
int64_t gspaln_h_task_cells(const gspaln_h_task* t);

#ifdef __cplusplus
}
#endif
#endif /* GSPALN_H */
