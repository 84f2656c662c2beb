#!/bin/bash
# Build-container tool (needs oracle/_ref, CPU only): the long oracle-vs-live-reference sweeps whose
# results are kept under profiles/r02_oracle_vs_reference_*.txt.  One process per option string (the
# reference keeps its options in globals).  About half an hour on one core.
# usage: tests/tools/run_live_sweeps.sh [output directory, default profiles]
cd "$(dirname "$0")/../.." || exit 1
D=${1:-profiles}
T=tests/tools
run() { # run <outfile> <label on a crash> <command...>
  local out=$1 label=$2; shift 2
  timeout 3000 "$@" > /tmp/live_sweep.$$ 2>&1; local rc=$?
  tail -1 /tmp/live_sweep.$$ >> "$out"
  [ $rc -gt 1 ] && echo "$label: rc=$rc" >> "$out"
  rm -f /tmp/live_sweep.$$
}

out=$D/r02_oracle_vs_reference_scalar_udh.txt
echo "# $T/sweep_oracle_scalar_udh.py: oracle restatements of the scalar Hirschberg passes (-A0) against the" > $out
echo "# unmodified reference (oracle/_ref) on random planted genes; one process per option string" >> $out
echo "# (the reference call runs in a forked child; passes on which the unmodified reference itself crashes are counted, not compared)" >> $out
for o in "-Q0 -A0 -yX0 -TDictyost" "-Q0 -A0 -yX0 -LS -TDictyost"; do
  for s in 11 12; do run $out "prot $o seed $s" python $T/sweep_oracle_scalar_udh.py prot 300 $s "$o"; done
done
for o in "-Q0 -A0 -S1 -yX0 -TDictyost" "-Q0 -A0 -S1 -yX0 -LS -TDictyost" "-Q0 -A0 -S1 -yX0 -yl3 -TDictyost" "-Q0 -A0 -S1 -yX0 -yl3 -LS -TDictyost"; do
  for s in 11 12; do run $out "dna $o seed $s" python $T/sweep_oracle_scalar_udh.py dna 400 $s "$o"; done
done

out=$D/r02_oracle_vs_reference_lsp_a0.txt
echo "# $T/sweep_oracle_lsp.py: driver restatements (so_lsp / so_lsp_h) against the unmodified reference's lspS_ng / lspH_ng" > $out
echo "# under -A0 (the default mode) on random planted genes; one process per option string" >> $out
for o in "-Q0 -A0 -S1 -yX0 -V64K -TDictyost" "-Q0 -A0 -S1 -yX0 -V64K -LS -TDictyost" "-Q0 -A0 -S1 -yX0 -V256K -yl3 -TDictyost" "-Q0 -A0 -S1 -yX0 -TDictyost"; do
  run $out "dna $o" python $T/sweep_oracle_lsp.py dna 400 21 "$o"
done
for o in "-Q0 -A0 -yX0 -V64K -TDictyost" "-Q0 -A0 -yX0 -V64K -LS -TDictyost" "-Q0 -A0 -yX0 -TDictyost"; do
  run $out "prot $o" python $T/sweep_oracle_lsp.py prot 300 21 "$o"
done

out=$D/r02_oracle_vs_reference_cip.txt
echo "# $T/sweep_oracle_lsp.py ... cip: driver restatements with Cip_score (annotated intron positions) against the" > $out
echo "# unmodified reference's lspS_ng / lspH_ng on random planted genes; one process per option string" >> $out
for o in "-Q0 -A2 -S1 -yX0 -V64K -TDictyost" "-Q0 -A2 -S1 -yX0 -LS -TDictyost" "-Q0 -A0 -S1 -yX0 -V64K -TDictyost" "-Q0 -A0 -S1 -yX0 -LS -TDictyost"; do
  run $out "dna $o" python $T/sweep_oracle_lsp.py dna 300 31 "$o" cip
done
for o in "-Q0 -A2 -yX0 -V64K -TDictyost" "-Q0 -A0 -yX0 -V64K -TDictyost" "-Q0 -A0 -yX0 -LS -TDictyost"; do
  run $out "prot $o" python $T/sweep_oracle_lsp.py prot 200 31 "$o" cip
done

out=$D/r02_oracle_vs_reference_recheck.txt
echo "# end of round 2: the round-1 sweeps once more after this round's edits to the oracle (sigB, windows, -A0 branches)" > $out
run $out scalar python $T/sweep_oracle_scalar.py 300 41 "-Q0 -A2 -S1 -yX0 -TDictyost"
run $out scalar python $T/sweep_oracle_scalar.py 300 42 "-Q0 -A2 -S1 -yX0 -LS -TDictyost"
run $out scalar python $T/sweep_oracle_scalar.py 300 43 "-Q0 -A2 -S1 -yX0 -yl3 -TDictyost"
run $out scalar_p python $T/sweep_oracle_scalar_p.py 200 41 "-Q0 -A2 -yX0 -TDictyost"
run $out scalar_p python $T/sweep_oracle_scalar_p.py 200 42 "-Q0 -A2 -yX0 -LS -TDictyost"
run $out protein_lsp python $T/sweep_oracle_protein_lsp.py 150 41
run $out protein_lsp python $T/sweep_oracle_protein_lsp.py -LS -V256K 150 42
run $out lsp python $T/sweep_oracle_lsp.py dna 300 41 "-Q0 -A2 -S1 -yX0 -V64K -TDictyost"
run $out lsp python $T/sweep_oracle_lsp.py dna 300 42 "-Q0 -A2 -S1 -yX0 -V256K -LS -TDictyost"
run $out lsp python $T/sweep_oracle_lsp.py dna 300 43 "-Q0 -A3 -S1 -yX0 -V64K -TDictyost"
run $out lsp python $T/sweep_oracle_lsp.py dna 200 44 "-Q0 -A6 -S1 -yX0 -V64K -TDictyost"
