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

}	// extern "C"
