// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI shim around the UNMODIFIED
// reference DNA spliced-DP translation unit.  The reference file is pulled in
// by path at compile time (-I/root/reference/src); nothing is copied.
//
// Why a textual include: `class Aln2s1` and the SIMD class `SimdAln2s1` are
// defined inside src/fwd2s1.cc (the latter via fwd2s1_simd.h / fwd2s1_simd.cc /
// fwd2s1_wip_simd.h, src/fwd2s1.cc:35-38) and their DP entry points are
// protected / not declared in any public header, so the only way to call
// `lspS_ng`, `forwardS1_wip`, `hirschbergS1_wip`, `scoreonlyS1_wip` exactly as
// `trcbkalignS_ng` (src/fwd2s1.cc:1667-1710) and `lspS_ng` (1801-1897) do is
// from inside that TU.

#include <chrono>
#include <cwchar>
#include "fwd2s1.cc"

namespace {

struct ShimAln2s1 : public Aln2s1 {
	ShimAln2s1(const Seq** s, const PwdB* p) : Aln2s1(s, p) {}
	// mirrors the head of Aln2s1::globalS_ng (src/fwd2s1.cc:2674-2683) with
	// algmode.qck == 0, but returns the raw Mfile SKL list (before
	// stdskl/trimskl) so the DP output itself can be compared.
	VTYPE	run_lsp(const WINDOW& wdw, SKL* out, int cap, int* n_out) {
	    mfd = new Mfile(sizeof(SKL));
	    VTYPE scr = lspS_ng(wdw);
	    int n = (int) mfd->size();
	    SKL* skl = (SKL*) mfd->flush();
	    *n_out = n;
	    for (int i = 0; i < n && i < cap; ++i) out[i] = skl[i];
	    delete[] skl;
	    delete mfd; mfd = 0;
	    return scr;
	}
	// Aln2s1::trcbkalignS_ng (src/fwd2s1.cc:1667-1710) with the SIMD level forced to 0, so that
	// its scalar branch (forwardS_ng + Vmf::traceback + end adjustment) runs for any number of
	// query rows -- the branch the stock code takes for blocks with fewer than 8 rows.
	VTYPE	run_scalar(const WINDOW& wdw, SKL* out, int cap, int* n_out) {
	    *const_cast<int*>(&simd) = 0;
	    mfd = new Mfile(sizeof(SKL));
	    VTYPE scr = trcbkalignS_ng(wdw);
	    int n = (int) mfd->size();
	    SKL* skl = (SKL*) mfd->flush();
	    *n_out = n;
	    for (int i = 0; i < n && i < cap; ++i) out[i] = skl[i];
	    delete[] skl;
	    delete mfd; mfd = 0;
	    return scr;
	}
	// Aln2s1::hirschbergS_ng (src/fwd2s1.cc:764-1104), the scalar Hirschberg pass of `-A0`, with the
	// spacing of the intermediate rows set as lspS_ng does (src/fwd2s1.cc:1840,1853)
	VTYPE	run_scalar_udh(const WINDOW& wdw, int n_imd, int intvl, Dim10* cpos) {
	    imd_intvl = intvl;
	    return hirschbergS_ng(cpos, n_imd, wdw);
	}
};

int copy_out(Mfile& mfd, SKL* out, int cap)
{
	int n = (int) mfd.size();
	SKL* skl = (SKL*) mfd.flush();
	for (int i = 0; i < n && i < cap; ++i) out[i] = skl[i];
	delete[] skl;
	return n;
}

}	// namespace

extern "C" {

// kind: 0 = forwardS1_wip (score + trace-back corners)
//       1 = scoreonlyS1_wip
//       2 = hirschbergS1_wip (cpos filled, n_imd intermediates)
// returns number of SKL corners written (kind 0), 0 otherwise.
int shim_s1_kernel(const Seq** seqs, const PwdB* pwd, int lw, int up,
	int kind, int n_imd, int mode, int* score, int* skl_out, int cap,
	int* cpos_out, double* seconds)
{
	WINDOW wdw = {lw, up, up - lw + 3};
	SpJunc spj(seqs[1], pwd);
	Mfile mfd(sizeof(SKL));
	int n = 0;
	auto t0 = std::chrono::steady_clock::now();
	if (kind == 0) {
	    SimdAln2s1 k(seqs, pwd, wdw, &spj, 0, 1, 0);
	    *score = k.forwardS1_wip(&mfd);
	} else if (kind == 1) {
	    SimdAln2s1 k(seqs, pwd, wdw, &spj, 0, 1, 0);
	    *score = k.scoreonlyS1_wip();
	} else {
	    Dim10* cpos = new Dim10[n_imd + 1];
	    for (int i = 0; i <= n_imd; ++i)
		cpos[i][0] = cpos[i][2] = end_of_ulk;
	    SimdAln2s1 k(seqs, pwd, wdw, &spj, 0, mode);
	    *score = k.hirschbergS1_wip(cpos, n_imd);
	    if (cpos_out)
		for (int i = 0; i <= n_imd; ++i)
		    for (int j = 0; j < 10; ++j) cpos_out[10 * i + j] = cpos[i][j];
	    delete[] cpos;
	}
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	if (kind == 0) n = copy_out(mfd, (SKL*) skl_out, cap);
	return n;
}

// the whole DP driver (trace-back vs multi-intermediate Hirschberg dispatch,
// src/fwd2s1.cc:1801-1897) on the current ranges of seqs[0], seqs[1]
int shim_s1_lsp(const Seq** seqs, const PwdB* pwd, int lw, int up,
	int* score, int* skl_out, int cap, double* seconds)
{
	WINDOW wdw = {lw, up, up - lw + 3};
	ShimAln2s1 alnv(seqs, pwd);
	int n = 0;
	auto t0 = std::chrono::steady_clock::now();
	*score = alnv.run_lsp(wdw, (SKL*) skl_out, cap, &n);
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	return n;
}

int shim_s1_nelem() { return Simd_functions<short>::Nelem; }

// the scalar trace-back kernel (Aln2s1::forwardS_ng through trcbkalignS_ng) on the current ranges
int shim_s1_scalar(const Seq** seqs, const PwdB* pwd, int lw, int up,
	int* score, int* skl_out, int cap, double* seconds)
{
	WINDOW wdw = {lw, up, up - lw + 3};
	ShimAln2s1 alnv(seqs, pwd);
	int n = 0;
	auto t0 = std::chrono::steady_clock::now();
	*score = alnv.run_scalar(wdw, (SKL*) skl_out, cap, &n);
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	return n;
}

// the scalar score-only kernel Aln2s1::scorealoneS_ng (src/fwd2s1.cc:1163-1336), what
// HomScoreS_ng runs under -A0 and for queries shorter than 4 residues (src/fwd2s1.cc:2704-2705)
int shim_s1_scorealone(const Seq** seqs, const PwdB* pwd, int lw, int up)
{
	WINDOW wdw = {lw, up, up - lw + 3};
	Aln2s1 alnv(seqs, pwd);
	return (int) alnv.scorealoneS_ng(wdw);
}

// the scalar Hirschberg pass Aln2s1::hirschbergS_ng on the current ranges; the Seq ranges are left
// as the pass narrowed them (read them back with ref_task_info)
int shim_s1_scalar_udh(const Seq** seqs, const PwdB* pwd, int lw, int up, int n_imd, int intvl,
	int* score, int* cpos_out)
{
	WINDOW wdw = {lw, up, up - lw + 3};
	ShimAln2s1 alnv(seqs, pwd);
	Dim10* cpos = new Dim10[n_imd + 1];
	for (int i = 0; i <= n_imd; ++i) {
	    for (int j = 0; j < 10; ++j) cpos[i][j] = 0;
	    cpos[i][0] = cpos[i][2] = end_of_ulk;
	}
	*score = (int) alnv.run_scalar_udh(wdw, n_imd, intvl, cpos);
	for (int i = 0; i <= n_imd; ++i)
	    for (int j = 0; j < 10; ++j) cpos_out[10 * i + j] = cpos[i][j];
	delete[] cpos;
	return 0;
}

}	// extern "C"
