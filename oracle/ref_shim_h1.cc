// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI shim around the UNMODIFIED reference protein x
// genome spliced-DP translation unit (src/fwd2h1.cc, which pulls in fwd2h1_simd.h and
// fwd2h1_wip_simd.h).  Same reasoning as ref_shim_s1.cc: `SimdAln2h1` lives inside that TU.
#include <chrono>
#include <cwchar>
#include "fwd2h1.cc"

namespace {
int copy_out_h(Mfile& mfd, SKL* out, int cap)
{
	int n = (int) mfd.size();
	SKL* skl = (SKL*) mfd.flush();
	for (int i = 0; i < n && i < cap; ++i) out[i] = skl[i];
	delete[] skl;
	return n;
}
}

extern "C" {

// kind 0: forwardH1_wip(mfd) (score + corners); kind 1: forwardH1_wip(0) (score only, as
// HomScoreH_ng calls it, src/fwd2h1.cc:3304-3306)
int shim_h1_kernel(const Seq** seqs, const PwdB* pwd, int lw, int up, int kind,
	int* score, int* skl_out, int cap, double* seconds)
{
	WINDOW wdw = {lw, up, up - lw + 7};
	SpJunc spj(seqs[1], pwd);
	Mfile mfd(sizeof(SKL));
	auto t0 = std::chrono::steady_clock::now();
	SimdAln2h1 k(seqs, pwd, wdw, &spj, 0, 1, 0);
	*score = k.forwardH1_wip(kind == 0? &mfd: 0);
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	return kind == 0? copy_out_h(mfd, (SKL*) skl_out, cap): 0;
}

// hirschbergH1_wip(cpos, n_imd) with the mode lspH_ng picks (src/fwd2h1.cc:2198-2207); cpos_out
// receives (n_imd + 1) x 10 ints.  The call narrows seqs[0/1]->left/right (the caller reads them).
int shim_h1_udh(const Seq** seqs, const PwdB* pwd, int lw, int up, int n_imd, int* score,
	int* cpos_out, double* seconds)
{
	WINDOW wdw = {lw, up, up - lw + 7};
	SpJunc spj(seqs[1], pwd);
	Dim10* cpos = new Dim10[n_imd + 1];
	for (int i = 0; i <= n_imd; ++i) vset(cpos[i], end_of_ulk, 10);
	auto t0 = std::chrono::steady_clock::now();
const	int mode = ((std::max(abs(wdw.lw), wdw.up) + wdw.width) < SHRT_MAX)? 2: 4;
	SimdAln2h1 k(seqs, pwd, wdw, &spj, 0, mode);
	*score = k.hirschbergH1_wip(cpos, n_imd);
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	for (int i = 0; i <= n_imd; ++i)
	    for (int j = 0; j < 10; ++j) cpos_out[10 * i + j] = cpos[i][j];
	delete[] cpos;
	return 0;
}

// the whole driver Aln2h1::lspH_ng (protected: reached through a derived class), returning the
// raw Mfile corner list
struct ShimAln2h1 : public Aln2h1 {
	ShimAln2h1(const Seq** sqs, const PwdB* pwd) : Aln2h1(sqs, pwd) {}
	VTYPE run_lsp(const WINDOW& w, SKL* out, int cap, int* n_out) {
	    mfd = new Mfile(sizeof(SKL));
	    VTYPE scr = lspH_ng(w);
	    int n = (int) mfd->size();
	    SKL* skl = (SKL*) mfd->flush();
	    *n_out = n;
	    for (int i = 0; i < n && i < cap; ++i) out[i] = skl[i];
	    delete[] skl;
	    delete mfd; mfd = 0;
	    return scr;
	}
};

// Aln2h1::trcbkalignH_ng (src/fwd2h1.cc:1997-2041) with the SIMD level forced to 0: its scalar branch
// (forwardH_ng + Vmf::traceback + end adjustment), what the stock code runs for blocks with < 8 rows
struct ShimAln2h1Scalar : public ShimAln2h1 {
	ShimAln2h1Scalar(const Seq** sqs, const PwdB* pwd) : ShimAln2h1(sqs, pwd) {}
	VTYPE run_scalar(const WINDOW& w, SKL* out, int cap, int* n_out) {
	    *const_cast<int*>(&simd) = 0;
	    mfd = new Mfile(sizeof(SKL));
	    VTYPE scr = trcbkalignH_ng(w, b->inex.intr);
	    int n = (int) mfd->size();
	    SKL* skl = (SKL*) mfd->flush();
	    *n_out = n;
	    for (int i = 0; i < n && i < cap; ++i) out[i] = skl[i];
	    delete[] skl;
	    delete mfd; mfd = 0;
	    return scr;
	}
};

// Aln2h1::hirschbergH_ng (src/fwd2h1.cc:1085-1520), the scalar Hirschberg pass of `-A0`, with the
// spacing of the intermediate rows set as lspH_ng does (src/fwd2h1.cc:2170,2183); the Seq ranges are
// left as the pass narrowed them
struct ShimAln2h1Udh : public ShimAln2h1 {
	ShimAln2h1Udh(const Seq** sqs, const PwdB* pwd) : ShimAln2h1(sqs, pwd) {}
	VTYPE run(const WINDOW& w, int n_imd, int intvl, Dim10* cpos) {
	    imd_intvl = intvl;
	    return hirschbergH_ng(cpos, n_imd, w);
	}
};

int shim_h1_scalar_udh(const Seq** seqs, const PwdB* pwd, int lw, int up, int n_imd, int intvl,
	int* score, int* cpos_out)
{
	WINDOW wdw = {lw, up, up - lw + 7};
	ShimAln2h1Udh aln(seqs, pwd);
	Dim10* cpos = new Dim10[n_imd + 1];
	for (int i = 0; i <= n_imd; ++i) vset(cpos[i], end_of_ulk, 10);
	*score = (int) aln.run(wdw, n_imd, intvl, cpos);
	for (int i = 0; i <= n_imd; ++i)
	    for (int j = 0; j < 10; ++j) cpos_out[10 * i + j] = cpos[i][j];
	delete[] cpos;
	return 0;
}

int shim_h1_scalar(const Seq** seqs, const PwdB* pwd, int lw, int up, int* score, int* skl_out, int cap)
{
	WINDOW wdw = {lw, up, up - lw + 7};
	int n = 0;
	ShimAln2h1Scalar aln(seqs, pwd);
	*score = aln.run_scalar(wdw, (SKL*) skl_out, cap, &n);
	return n;
}

// split-codon tables of SpJunc::spjseq (src/codepot.h:130-190) and aa2nuc (src/seq.cc:76):
// out = spj_tron_tab[257][2] | spj_amb_tron_tab[64][2] | spj_tron_amb_tab[64][2] | aa2nuc[26]
void shim_h1_spj_tables(unsigned char* out)
{
	int k = 0;
	for (int i = 0; i < 257; ++i) { out[k++] = spj_tron_tab[i][0]; out[k++] = spj_tron_tab[i][1]; }
	for (int i = 0; i < 64; ++i) { out[k++] = spj_amb_tron_tab[i][0]; out[k++] = spj_amb_tron_tab[i][1]; }
	for (int i = 0; i < 64; ++i) { out[k++] = spj_tron_amb_tab[i][0]; out[k++] = spj_tron_amb_tab[i][1]; }
	for (int i = 0; i < 26; ++i) out[k++] = aa2nuc[i];
}

int shim_h1_lsp(const Seq** seqs, const PwdB* pwd, int lw, int up, int* score, int* skl_out,
	int cap, double* seconds)
{
	WINDOW wdw = {lw, up, up - lw + 7};
	int n = 0;
	auto t0 = std::chrono::steady_clock::now();
	ShimAln2h1 aln(seqs, pwd);
	*score = aln.run_lsp(wdw, (SKL*) skl_out, cap, &n);
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	return n;
}

}	// extern "C"
