// TEST INFRASTRUCTURE ONLY (oracle/): runs include/gspaln_spaln_adapter.hpp -- the adapter a
// Spaln maintainer would add -- on the reference's own Seq / PwdB / Exinon objects, which live in
// libspaln_ref.so.  Built as its own library (oracle/_ref/libspaln_dropin.so, links libgspaln) so
// that the reference checker itself never maps the product.  The reference headers come from
// -I/root/reference/src at compile time; nothing is copied.
#include <cwchar>
#include "aln.h"
#include "vmf.h"
#include "gsinfo.h"
#include "gspaln_spaln_adapter.hpp"

namespace {
int copy_out(Mfile& mfd, SKL* out, int cap)
{
	int n = (int) mfd.size();
	SKL* skl = (SKL*) mfd.flush();
	for (int i = 0; i < n && i < cap; ++i) out[i] = skl[i];
	delete[] skl;
	return n;
}
}

extern "C" {

// SimdAln2s1::forwardS1_wip (kind 0) / scoreonlyS1_wip (kind 1) through the adapter (GPU)
int dropin_s1_adapter(const Seq** seqs, const PwdB* pwd, int lw, int up,
	int kind, int device, int* score, int* skl_out, int cap)
{
	static gspaln::SpalnEngine* eng = 0;
	if (!eng) eng = new gspaln::SpalnEngine(pwd, device, seqs[1]->inex.intr);
	WINDOW wdw = {lw, up, up - lw + 3};
	if (kind == 1) {
	    *score = eng->scoreonlyS1_wip(seqs, wdw);
	    return 0;
	}
	Mfile mfd(sizeof(SKL));
	*score = eng->forwardS1_wip(seqs, wdw, &mfd);
	return copy_out(mfd, (SKL*) skl_out, cap);
}

// Aln2s1::lspS_ng through the adapter (GPU): int53 / sig53tab come from the caller because they
// are private to Exinon (the reference shim exports the pointers).  Returns the number of
// corners, or -1 if the adapter reports the problem as unsupported.
int dropin_s1_adapter_lsp(const Seq** seqs, const PwdB* pwd, int lw, int up, int device,
	const void* int53, const short* sig53tab, int* score, int* skl_out, int cap)
{
	static gspaln::SpalnEngine* eng = 0;
	if (!eng) {
	    eng = new gspaln::SpalnEngine(pwd, device, seqs[1]->inex.intr);
	    if (sig53tab) eng->enable_scalar(pwd, sig53tab, 1 << 19);
	}
	WINDOW wdw = {lw, up, up - lw + 3};
	Mfile mfd(sizeof(SKL));
	VTYPE scr = 0;
	// the Aln2s1 constructor's Cip_score for an annotated query (src/fwd2s1.cc:124)
	const Cip_score* cip = (seqs[1]->inex.intr && seqs[0]->sigII)? new Cip_score(seqs[0]): 0;
	const bool ok = eng->lspS_ng(seqs, wdw, &mfd, (const INT53*) int53, &scr, cip);
	delete cip;
	if (!ok) return -1;
	*score = scr;
	return copy_out(mfd, (SKL*) skl_out, cap);
}

// SimdAln2h1::forwardH1_wip(mfd) (kind 0) / forwardH1_wip(0) (kind 1) through the adapter
int dropin_h1_adapter(const Seq** seqs, const PwdB* pwd, int lw, int up,
	int kind, int device, int* score, int* skl_out, int cap)
{
	static gspaln::SpalnEngineH* eng = 0;
	if (!eng) eng = new gspaln::SpalnEngineH(pwd, device, seqs[1]->inex.intr);
	WINDOW wdw = {lw, up, up - lw + 7};
	if (kind == 1) {
	    *score = eng->forwardH1_wip(seqs, wdw, 0);
	    return 0;
	}
	Mfile mfd(sizeof(SKL));
	*score = eng->forwardH1_wip(seqs, wdw, &mfd);
	return copy_out(mfd, (SKL*) skl_out, cap);
}

// Aln2h1::lspH_ng through the adapter; spj_tabs: the 796 bytes of gspaln_h_set_ng_tables
int dropin_h1_adapter_lsp(const Seq** seqs, const PwdB* pwd, int lw, int up, int device,
	const void* int53, const short* sig53tab, const unsigned char* spj_tabs,
	int* score, int* skl_out, int cap)
{
	static gspaln::SpalnEngineH* eng = 0;
	if (!eng) {
	    eng = new gspaln::SpalnEngineH(pwd, device, seqs[1]->inex.intr);
	    if (sig53tab && spj_tabs) eng->enable_scalar(pwd, sig53tab, spj_tabs, 1 << 19);
	}
	WINDOW wdw = {lw, up, up - lw + 7};
	Mfile mfd(sizeof(SKL));
	VTYPE scr = 0;
	// the Aln2h1 constructor's Cip_score for an annotated query (src/fwd2h1.cc:126)
	const Cip_score* cip = (seqs[1]->inex.intr && seqs[0]->sigII)? new Cip_score(seqs[0]): 0;
	const bool ok = eng->lspH_ng(seqs, wdw, &mfd, (const INT53*) int53, &scr, cip);
	delete cip;
	if (!ok) return -1;
	*score = scr;
	return copy_out(mfd, (SKL*) skl_out, cap);
}

}	// extern "C"
