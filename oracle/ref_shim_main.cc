// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI shim that sets the reference up
// exactly the way its own `main` does and exposes per-task objects so that
// tests / the CPU-baseline leg can (1) harvest the raw inputs the DP kernels
// consume (sequence codes, SGPT2 splice-signal table, frozen parameters) and
// (2) run the reference's own kernels on them.
//
// The reference's start-up code (option parser, defaults, SetUpPwd) is a set
// of file-static functions in src/spaln.cc, so this TU includes that file by
// path (never copied) with `main` renamed.  See oracle/Makefile.

// system headers first: the reference overloads fclose() for gzFile, which
// breaks glibc attribute checks if <wchar.h> is first seen afterwards
#include <string>
#include <vector>
#include <cwchar>

#define main spaln_reference_main_unused
#include "spaln.cc"
#undef main

// Exinon keeps its INT53 array and sig53tab private (src/codepot.h:72-80); the harness exports
// them as inputs of the scalar kernel.  Explicit instantiation may name private members, which
// gives access without touching the reference headers.
namespace {
template <typename Tag, typename Tag::type M> struct Rob {
	friend typename Tag::type get(Tag) { return M; }
};
struct ExinonInt53 { typedef INT53* Exinon::*type; friend type get(ExinonInt53); };
struct ExinonTab { typedef STYPE** Exinon::*type; friend type get(ExinonTab); };
template struct Rob<ExinonInt53, &Exinon::int53>;
template struct Rob<ExinonTab, &Exinon::sig53tab>;
}

extern "C" {
int shim_s1_kernel(const Seq** seqs, const PwdB* pwd, int lw, int up,
	int kind, int n_imd, int mode, int* score, int* skl_out, int cap,
	int* cpos_out, double* seconds);
int shim_s1_lsp(const Seq** seqs, const PwdB* pwd, int lw, int up,
	int* score, int* skl_out, int cap, double* seconds);
int shim_s1_nelem();
int shim_s1_scorealone(const Seq** seqs, const PwdB* pwd, int lw, int up);
int shim_s1_scalar(const Seq** seqs, const PwdB* pwd, int lw, int up,
	int* score, int* skl_out, int cap, double* seconds);
int shim_s1_scalar_udh(const Seq** seqs, const PwdB* pwd, int lw, int up, int n_imd, int intvl,
	int* score, int* cpos_out);
int shim_h1_udh(const Seq** seqs, const PwdB* pwd, int lw, int up, int n_imd, int* score,
	int* cpos_out, double* seconds);
int shim_h1_lsp(const Seq** seqs, const PwdB* pwd, int lw, int up, int* score, int* skl_out,
	int cap, double* seconds);
int shim_h1_scalar(const Seq** seqs, const PwdB* pwd, int lw, int up, int* score, int* skl_out, int cap);
int shim_h1_scalar_udh(const Seq** seqs, const PwdB* pwd, int lw, int up, int n_imd, int intvl,
	int* score, int* cpos_out);
void shim_h1_spj_tables(unsigned char* out);
int shim_h1_kernel(const Seq** seqs, const PwdB* pwd, int lw, int up, int kind,
	int* score, int* skl_out, int cap, double* seconds);
}

namespace {
PwdB*	g_pwd = 0;
std::vector<std::string> g_args;

struct RefTask {
	Seq*	sqs[4];
};
}

extern "C" {

// optstr: the spaln command-line options, blank separated, e.g.
// "-Q0 -A2 -S1 -yX0 -TDictyost".  genome_fa / query_fa: any DNA FASTA pair,
// only used for molecule-type inference exactly as main() does
// (src/spaln.cc:1496-1501,1588-1603).
int ref_setup(const char* optstr, const char* genome_fa, const char* query_fa)
{
	if (g_pwd) return 1;
	g_args.clear();
	g_args.push_back("spaln");
	std::string s(optstr? optstr: "");
	size_t p = 0;
	while (p < s.size()) {
	    while (p < s.size() && isspace(s[p])) ++p;
	    size_t q = p;
	    while (q < s.size() && !isspace(s[q])) ++q;
	    if (q > p) g_args.push_back(s.substr(p, q - p));
	    p = q;
	}
	g_args.push_back(genome_fa);
	g_args.push_back(query_fa);
	std::vector<const char*> argv;
	for (auto& a : g_args) argv.push_back(a.c_str());
	int argc = (int) argv.size();

	setdefparam();
	optimize(GLOBAL, MAXIMUM);
	int n = getoption(argc, argv.data());
	spb_fact();
	(void) n;
	thread_num = 0;
	if (!algmode.mlt) OutPrm.MaxOut = 1;
	if (OutPrm.MaxOut > OutPrm.MaxOut2) OutPrm.MaxOut2 = OutPrm.MaxOut;
	Seq*	seqs[3];
	initseq(seqs, 3);
	if (!seqs[1]->getseq(genome_fa)) return -1;
	if (!seqs[0]->getseq(query_fa)) return -2;
	g_pwd = SetUpPwd(seqs);
	// Aln2s1's constructor flattens the ILD under -A3 (src/fwd2s1.cc:125);
	// kernel-level calls bypass that constructor, so apply it here
	if ((algmode.alg & 3) == 3) IntronPrm.nquant = 1;
	b_intr = seqs[1]->inex.intr;
	if (seqs[0]->inex.intr || seqs[1]->inex.intr) makeStdSig53();
	clearseq(seqs, 3);
	return 0;
}

int ref_nelem() { return shim_s1_nelem(); }

// frozen parameter block (all int32).  Layout documented in
// tests/ref_harness.py::PARAM_FIELDS.
int ref_get_params(int* out, int cap, int* simmtx_out, int simmtx_cap)
{
	if (!g_pwd) return -1;
	int k = 0;
	const PwdB* p = g_pwd;
	int vals[] = {
	    p->DvsP, p->Noll, (int) p->Vab, (int) p->Vthr,
	    (int) p->BasicGOP, (int) p->BasicGEP, (int) p->LongGOP, (int) p->LongGEP,
	    p->codonk1,
	    p->IntPen? (int) p->IntPen->Penalty(): 0,
	    IntronPrm.llmt, IntronPrm.mu, IntronPrm.rlmt, IntronPrm.minl,
	    IntronPrm.maxl, IntronPrm.mode, IntronPrm.nquant,
	    (int) algmode.alg, (int) algmode.lcl, (int) algmode.lsg,
	    (int) algmode.any, (int) algmode.qck,
	    (int) alprm.sh, (int) alprm.ubh, (int) alprm.scale,
	    (int) p->simmtx->AvTrc(), p->simmtx->dim, MaxVmfSpace,
	    (int) p->GapPenalty(1), ref_nelem()
	};
	int nv = sizeof(vals) / sizeof(int);
	for (int i = 0; i < nv && k < cap; ++i) out[k++] = vals[i];
	// quantile table (len, pen) x nquant
	for (int j = 0; j < IntronPrm.nquant && p->IntPen && p->IntPen->qm; ++j) {
	    if (k + 2 > cap) break;
	    out[k++] = p->IntPen->qm[j].len;
	    out[k++] = p->IntPen->qm[j].pen;
	}
	int d = p->simmtx->dim;
	if (simmtx_out && simmtx_cap >= d * d)
	    for (int i = 0; i < d; ++i)
		for (int j = 0; j < d; ++j)
		    simmtx_out[i * d + j] = p->simmtx->mtx[i][j];
	return k;
}

// build the (query, genome) pair the way match_2 does for cDNA x genome
// (src/spaln.cc:734-765): Exinon on b, end-gap flags from algmode.lcl.
void* ref_task_new(const char* genome_fa, const char* query_fa, int comrev_query)
{
	if (!g_pwd) return 0;
	RefTask* t = new RefTask;
	initseq(t->sqs, 4);
	Seq*& a = t->sqs[0];
	Seq*& b = t->sqs[1];
	if (!b->getseq(genome_fa) || !a->getseq(query_fa)) {
	    clearseq(t->sqs, 4);
	    delete t;
	    return 0;
	}
	a->inex.intr = 0;
	b->inex.intr = b_intr;
	if (comrev_query) a->comrev();
	if (g_pwd->DvsP == 1) b->nuc2tron();		// protein query: src/spaln.cc:748-753
	if (!b->exin) b->exin = new Exinon(b, g_pwd, false);
	if (algmode.lcl & 16) {
	    a->exg_seq(1, 1);
	    b->exg_seq(1, 1);
	} else {
	    a->exg_seq(algmode.lcl & 4, algmode.lcl & 8);
	    b->exg_seq(algmode.lcl & 1, algmode.lcl & 2);
	}
	return t;
}

void ref_task_free(void* h)
{
	RefTask* t = (RefTask*) h;
	if (!t) return;
	clearseq(t->sqs, 4);
	delete t;
}

// info[0..9] = a.len, b.len, a.left, a.right, b.left, b.right,
//              a.exgl, a.exgr, b.exgl, b.exgr
void ref_task_info(void* h, int* info)
{
	RefTask* t = (RefTask*) h;
	const Seq* a = t->sqs[0];
	const Seq* b = t->sqs[1];
	info[0] = a->len; info[1] = b->len;
	info[2] = a->left; info[3] = a->right;
	info[4] = b->left; info[5] = b->right;
	info[6] = a->inex.exgl; info[7] = a->inex.exgr;
	info[8] = b->inex.exgl; info[9] = b->inex.exgr;
}

void ref_task_set(void* h, const int* info)
{
	RefTask* t = (RefTask*) h;
	const Seq* a = t->sqs[0];
	const Seq* b = t->sqs[1];
	a->left = info[2]; a->right = info[3];
	b->left = info[4]; b->right = info[5];
	a->inex.exgl = info[6]; a->inex.exgr = info[7];
	b->inex.exgl = info[8]; b->inex.exgr = info[9];
}

// a_codes[a.len + 2], b_codes[b.len + 2]: residue codes at(-1 .. len)
// sig5/sig3[b.len + 2]: SGPT2 table entries for columns n = 0 .. b.len + 1
void ref_task_export(void* h, unsigned char* a_codes, unsigned char* b_codes,
	short* sig5, short* sig3)
{
	RefTask* t = (RefTask*) h;
	const Seq* a = t->sqs[0];
	const Seq* b = t->sqs[1];
	for (int i = -1; i <= a->len; ++i) a_codes[i + 1] = *a->at(i);
	for (int i = -1; i <= b->len; ++i) b_codes[i + 1] = *b->at(i);
	for (int n = 0; n <= b->len + 1; ++n) {
	    const SGPT2* g = b->exin->score_n(n);
	    sig5[n] = g->sig5;
	    sig3[n] = g->sig3;
	}
}

// overwrite the splice-signal table (synthetic workloads)
void ref_task_inject(void* h, const short* sig5, const short* sig3)
{
	RefTask* t = (RefTask*) h;
	const Seq* b = t->sqs[1];
	for (int n = 0; n <= b->len + 1; ++n) {
	    SGPT2* g = b->exin->score_n(n);
	    g->sig5 = sig5[n];
	    g->sig3 = sig3[n];
	}
}

// protein x genome: SGPT6 table (8 shorts per column n = 0 .. b.len + 1:
// sig5, sig3, sigS, sigT, sigE, sigI, phs5, phs3) and the frozen protein parameters
void ref_task_export_p(void* h, unsigned char* a_codes, unsigned char* b_codes, short* sgpt6)
{
	RefTask* t = (RefTask*) h;
	const Seq* a = t->sqs[0];
	const Seq* b = t->sqs[1];
	for (int i = -1; i <= a->len; ++i) a_codes[i + 1] = *a->at(i);
	for (int i = -1; i <= b->len; ++i) b_codes[i + 1] = *b->at(i);
	for (int n = 0; n <= b->len + 1; ++n) {
	    const SGPT6* g = b->exin->score_p(n);
	    short* o = sgpt6 + 8 * n;
	    o[0] = g->sig5; o[1] = g->sig3; o[2] = g->sigS; o[3] = g->sigT;
	    o[4] = g->sigE; o[5] = g->sigI; o[6] = g->phs5; o[7] = g->phs3;
	}
}

// overwrite the SGPT6 table (same 8-shorts-per-column layout as ref_task_export_p): lets both
// arms of a benchmark run on one synthetic table
void ref_task_inject_p(void* h, const short* sgpt6)
{
	RefTask* t = (RefTask*) h;
	const Seq* b = t->sqs[1];
	for (int n = 0; n <= b->len + 1; ++n) {
	    SGPT6* g = b->exin->score_p(n);
	    const short* o = sgpt6 + 8 * n;
	    g->sig5 = o[0]; g->sig3 = o[1]; g->sigS = o[2]; g->sigT = o[3];
	    g->sigE = o[4]; g->sigI = o[5]; g->phs5 = (char) o[6]; g->phs3 = (char) o[7];
	}
}

int ref_get_params_p(int* out, int cap)
{
	if (!g_pwd) return -1;
	const PwdB* p = g_pwd;
	int vals[] = {p->GapW1, p->GapW2, p->GapW3, p->GapW3L, p->GapE1, p->GapE2, p->ExtraGOP,
	    p->codonk1, (int) alprm2.termk1, p->simmtx->rows, p->simmtx->cols};
	int nv = sizeof(vals) / sizeof(int);
	for (int i = 0; i < nv && i < cap; ++i) out[i] = vals[i];
	return nv;
}

int ref_task_kernel_p(void* h, int lw, int up, int kind, int* score, int* skl_out, int cap,
	double* seconds)
{
	RefTask* t = (RefTask*) h;
	return shim_h1_kernel((const Seq**) t->sqs, g_pwd, lw, up, kind, score, skl_out, cap, seconds);
}

int ref_task_udh_p(void* h, int lw, int up, int n_imd, int* score, int* cpos_out, double* seconds)
{
	RefTask* t = (RefTask*) h;
	return shim_h1_udh((const Seq**) t->sqs, g_pwd, lw, up, n_imd, score, cpos_out, seconds);
}

int ref_task_lsp_p(void* h, int lw, int up, int* score, int* skl_out, int cap, double* seconds)
{
	RefTask* t = (RefTask*) h;
	return shim_h1_lsp((const Seq**) t->sqs, g_pwd, lw, up, score, skl_out, cap, seconds);
}

void ref_task_stripe31(void* h, int sh, int* lwup)
{
	RefTask* t = (RefTask*) h;
	WINDOW w;
	stripe31((const Seq**) t->sqs, &w, sh);
	lwup[0] = w.lw; lwup[1] = w.up; lwup[2] = w.width;
}

void ref_task_stripe(void* h, int sh, int* lwup)
{
	RefTask* t = (RefTask*) h;
	WINDOW w;
	stripe((const Seq**) t->sqs, &w, sh);
	lwup[0] = w.lw; lwup[1] = w.up; lwup[2] = w.width;
}

int ref_task_kernel(void* h, int lw, int up, int kind, int n_imd, int mode,
	int* score, int* skl_out, int cap, int* cpos_out, double* seconds)
{
	RefTask* t = (RefTask*) h;
	return shim_s1_kernel((const Seq**) t->sqs, g_pwd, lw, up, kind, n_imd,
	    mode, score, skl_out, cap, cpos_out, seconds);
}

// raw handles for oracle/dropin_shim.cc (libspaln_dropin.so), which runs the same Seq / PwdB
// objects through include/gspaln_spaln_adapter.hpp
// Attach intron-position annotation to the query (what a `;B` / `;b` block of the query file
// produces, src/gsinfo.h:76-126): n boundaries at positions pos[] with multiplicities num[].
// The Aln2* constructors build their Cip_score from it (src/fwd2s1.cc:124, src/fwd2h1.cc:126).
void ref_task_set_cip(void* h, const int* pos, const int* num, int n)
{
	Seq* a = ((RefTask*) h)->sqs[0];
	delete a->sigII;
	a->sigII = new SigII(pos, n, a->isprotein()? 3: 1);
	for (int i = 0; i < n; ++i) {
	    a->sigII->pfq[i].num = num[i];
#if USE_WEIGHT
	    a->sigII->pfq[i].dns = num[i];
#endif
	}
	if (n) a->sigII->pfq[n] = pfqend;
}

// Cip_score::cip_score(c) for c in [0, n) as the DP kernels of this task would see it
void ref_task_cip_table(void* h, int* out, int n)
{
	Cip_score cs(((RefTask*) h)->sqs[0]);
	for (int c = 0; c < n; ++c) out[c] = (int) cs.cip_score(c);
}

// Aln2s1::hirschbergS_ng (the scalar Hirschberg pass of `-A0`) on the task's current ranges
int ref_task_scalar_udh(void* h, int lw, int up, int n_imd, int intvl, int* score, int* cpos_out)
{
	RefTask* t = (RefTask*) h;
	return shim_s1_scalar_udh((const Seq**) t->sqs, g_pwd, lw, up, n_imd, intvl, score, cpos_out);
}

// Aln2h1::hirschbergH_ng (the scalar protein Hirschberg pass of `-A0`) on the task's current ranges
int ref_task_scalar_udh_p(void* h, int lw, int up, int n_imd, int intvl, int* score, int* cpos_out)
{
	RefTask* t = (RefTask*) h;
	return shim_h1_scalar_udh((const Seq**) t->sqs, g_pwd, lw, up, n_imd, intvl, score, cpos_out);
}

const void* ref_pwd() { return g_pwd; }
void* ref_task_seqs(void* h) { return ((RefTask*) h)->sqs; }
const void* ref_task_int53_ptr(void* h)
{
	const Seq* b = ((RefTask*) h)->sqs[1];
	return b->exin? (const void*) (b->exin->*get(ExinonInt53())): 0;
}
const void* ref_task_sig53tab_ptr(void* h)
{
	const Seq* b = ((RefTask*) h)->sqs[1];
	if (!b->exin) return 0;
	STYPE** tab = b->exin->*get(ExinonTab());
	return tab? (const void*) tab[0]: 0;
}

// inputs of the exact intron scoring (Aln2s1::forwardS_ng): per column n = 0 .. b.len + 1 the
// INT53 nibbles (dinc5 | dinc3 << 4 | cano5 << 8 | cano3 << 12, src/codepot.h:49-54)
void ref_task_export_int53(void* h, unsigned short* out)
{
	RefTask* t = (RefTask*) h;
	const Seq* b = t->sqs[1];
	for (int n = 0; n <= b->len + 1; ++n) {
	    const INT53& w = (b->exin->*get(ExinonInt53()))[n];
	    out[n] = (unsigned short) (w.dinc5 | (w.dinc3 << 4) | (w.cano5 << 8) | (w.cano3 << 12));
	}
}

// sig53tab (544 shorts, src/codepot.cc:281-285) of this task's Exinon and
// IntronPenalty::Penalty(n) for n in [0, n_pen); misc[0] = alprm2.Z > 0
int ref_task_export_ng_tables(void* h, short* sig53tab, short* penalty, int n_pen, int* misc)
{
	RefTask* t = (RefTask*) h;
	const Seq* b = t->sqs[1];
	if (!b->exin || !g_pwd->IntPen) return -1;
	STYPE** tab = b->exin->*get(ExinonTab());
	if (!tab) return -1;
	for (int i = 0; i < 544; ++i) sig53tab[i] = tab[0][i];
	for (int n = 0; n < n_pen; ++n) penalty[n] = g_pwd->IntPen->Penalty(n);
	misc[0] = alprm2.Z > 0;
	return 0;
}

// splice-site PSSMs (EijPat::pattern5 / pattern3, src/utilseq.h:169-180) and the factors
// Exinon::intron53_n applies to them (src/codepot.cc:479-523): inputs of the signal scan.
// which: 0 = pattern5, 1 = pattern3.  meta = rows, cols, offset, nalpha, morder;
// fmeta = tonic, min_elem.  Returns the number of matrix floats (rows * cols), 0 if absent.
int ref_get_patmat(int which, int* meta, float* fmeta, float* mtx, int cap)
{
	if (!g_pwd || !g_pwd->eijpat) return -1;
	const EijPat* ep = g_pwd->eijpat;
	const PatMat* pm = which == 0? ep->pattern5: which == 1? ep->pattern3:
	    which == 2? ep->patternI: which == 3? ep->patternT: ep->patternB;
	if (!pm) return 0;
	meta[0] = pm->rows; meta[1] = pm->cols; meta[2] = pm->offset;
	meta[3] = pm->nalpha; meta[4] = pm->order();
	fmeta[0] = pm->tonic; fmeta[1] = pm->min_elem;
	int n = pm->rows * pm->cols;
	for (int i = 0; i < n && i < cap; ++i) mtx[i] = pm->mtx[i];
	return n;
}

// fvals = Exinon::fS of this task, alprm2.sss, EijPat::tonic5, tonic3; ivals = algmode.any
void ref_task_scan_factors(void* h, float* fvals, int* ivals)
{
	RefTask* t = (RefTask*) h;
	const Seq* b = t->sqs[1];
	fvals[0] = b->exin->fS; fvals[1] = alprm2.sss;
	fvals[2] = g_pwd->eijpat->tonic5; fvals[3] = g_pwd->eijpat->tonic3;
	ivals[0] = algmode.any;
	ivals[1] = b->inex.cmpc;
	ivals[2] = b->many;
}

// the genetic code table Seq::nuc2tron translates with (src/utilseq.cc:38, set by -C)
void ref_get_gencode(unsigned char* out64)
{
	for (int i = 0; i < 64; ++i) out64[i] = gencode[i];
}

// protein-side scan (Exinon::intron53_p, src/codepot.cc:525-619): which potentials exist and the
// factors applied to them.  ivals: codepot present, its size() (ndata), dsize(), exonpot present,
// intnpot present, DvsP, bpprm.maxb3d; fvals: Exinon::fact, alprm2.z, alprm2.Z, alprm2.bti,
// bpprm.factor, alprm2.o, EijPat::tonicB
void ref_task_scan_factors_p(void* h, float* fvals, int* ivals)
{
	RefTask* t = (RefTask*) h;
	const Seq* b = t->sqs[1];
	ivals[0] = g_pwd->codepot != 0;
	ivals[1] = g_pwd->codepot? g_pwd->codepot->size(): 0;
	ivals[2] = g_pwd->codepot? g_pwd->codepot->dsize(): 0;
	ivals[3] = g_pwd->exonpot != 0;
	ivals[4] = g_pwd->intnpot != 0;
	ivals[5] = g_pwd->DvsP;
	ivals[6] = bpprm.maxb3d;
	fvals[0] = b->exin->fact; fvals[1] = alprm2.z; fvals[2] = alprm2.Z; fvals[3] = alprm2.bti;
	fvals[4] = bpprm.factor; fvals[5] = alprm2.o; fvals[6] = g_pwd->eijpat->tonicB;
}

// the coding potential table (ExinPot::begin(), dsize() floats)
int ref_get_codepot(float* out, int cap)
{
	if (!g_pwd || !g_pwd->codepot) return 0;
	int n = g_pwd->codepot->dsize();
	for (int i = 0; i < n && i < cap; ++i) out[i] = g_pwd->codepot->begin()[i];
	return n;
}

int ref_task_scalar_p(void* h, int lw, int up, int* score, int* skl_out, int cap)
{
	RefTask* t = (RefTask*) h;
	return shim_h1_scalar((const Seq**) t->sqs, g_pwd, lw, up, score, skl_out, cap);
}

// spj tables (see shim_h1_spj_tables) and the protein-side scalar parameters:
// ivals = IntronPrm.minl, ExtraGOP, GapW3L, alprm2.termk1
void ref_get_scalar_p(unsigned char* tabs, int* ivals)
{
	shim_h1_spj_tables(tabs);
	ivals[0] = IntronPrm.minl; ivals[1] = g_pwd->ExtraGOP; ivals[2] = g_pwd->GapW3L;
	ivals[3] = (int) alprm2.termk1;
}

int ref_task_scorealone(void* h, int lw, int up)
{
	RefTask* t = (RefTask*) h;
	return shim_s1_scorealone((const Seq**) t->sqs, g_pwd, lw, up);
}

int ref_task_scalar(void* h, int lw, int up, int* score, int* skl_out, int cap, double* seconds)
{
	RefTask* t = (RefTask*) h;
	return shim_s1_scalar((const Seq**) t->sqs, g_pwd, lw, up, score, skl_out, cap, seconds);
}


int ref_task_lsp(void* h, int lw, int up, int* score, int* skl_out, int cap,
	double* seconds)
{
	RefTask* t = (RefTask*) h;
	return shim_s1_lsp((const Seq**) t->sqs, g_pwd, lw, up, score, skl_out,
	    cap, seconds);
}

}	// extern "C"
