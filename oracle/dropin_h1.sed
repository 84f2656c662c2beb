# the INTEGRATION.md patch of src/fwd2h1.cc: one #include and three one-line hooks
/^class Aln2h1 {/i #include "gspaln_spaln_dropin.hpp"
/^VTYPE Aln2h1::lspH_ng(const WINDOW& wdw)$/{n;s/^{$/{ GSPALN_HOOK_LSPH/}
/^VTYPE Aln2h1::trcbkalignH_ng(const WINDOW& wdw, bool spj, const RANGE\* mc)$/{n;s/^{$/{ GSPALN_HOOK_TRCBKH/}
/^VTYPE HomScoreH_ng(const Seq\* seqs\[\], const PwdB\* pwd)$/{n;s/^{$/{ GSPALN_HOOK_HOMH/}
