# the INTEGRATION.md patch of src/fwd2s1.cc: one #include and three one-line hooks
/^class Aln2s1 {/i #include "gspaln_spaln_dropin.hpp"
/^VTYPE Aln2s1::lspS_ng(const WINDOW& wdw)$/{n;s/^{$/{ GSPALN_HOOK_LSPS/}
/^VTYPE Aln2s1::trcbkalignS_ng(const WINDOW& wdw, bool spj, const RANGE\* mc)$/{n;s/^{$/{ GSPALN_HOOK_TRCBKS/}
/^VTYPE HomScoreS_ng(const Seq\* seqs\[\], const PwdB\* pwd)$/{n;s/^{$/{ GSPALN_HOOK_HOMS/}
