#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle).  Regenerates tests/golden/*: outputs of the UNMODIFIED reference (oracle/_ref/deSAMBA_zero =
# reference sources + -ftrivial-auto-var-init=zero, run -t 1: the parity oracle O_def of SURVEY.md 8c; and deSAMBA_stock -t 4
# for the demo) on the demo reads and on small deterministic synthetic sets (desamba_b200/bin/simreads, seeds below).
# Needs oracle/_ref (oracle/build_ref.sh + oracle/build_index.sh, i.e. /root/reference present).  Outputs are gzip'd text.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(dirname "$HERE")
R=$HERE/_ref; G=$ROOT/tests/golden; SIM=$ROOT/desamba_b200/bin/simreads
IDX=$R/demo/idx; FA=$R/demo/viral-gs.fa
mkdir -p "$G" "$R/sets"
[ -s "$IDX/deSAMBA.bwt" ] || "$HERE/build_index.sh" "$FA" "$IDX"
# name mode n err seed
gen() { [ -s "$R/sets/$1.fq" ] || "$SIM" "$2" "$FA" "$3" "$4" "$5" "$R/sets/$1.fq"; }
gen long10  long  300  0.10 20261020
gen long30  long  300  0.30 20261021
gen short1  short 5000 0.01 20261022
[ -s "$R/sets/mixed.fq" ] || "$SIM" mixed "$FA" 150 1500 20261024 "$R/sets/mixed.fq"
run() { # set fmt extra-opts tag
  "$R/deSAMBA_zero" classify -t 1 -f "$2" $3 "$IDX" "$R/sets/$1.fq" -o "$G/$1$4.$2" 2>/dev/null; gzip -9nf "$G/$1$4.$2"; }
for s in long10 long30 short1 mixed; do run $s DES_FULL "" ""; run $s SAM "" ""; done
run long10 SAM "-l 100 -s 40 -r 2" ".l100s40r2"
for f in SAM SAM_FULL DES DES_FULL; do
  "$R/deSAMBA_stock" classify -t 4 -f $f "$IDX" "$R/demo/ERR1050068.fastq" -o "$G/demo.$f" 2>/dev/null
  "$R/deSAMBA_zero" classify -t 1 -f $f "$IDX" "$R/demo/ERR1050068.fastq" -o "$G/demo.zero.$f" 2>/dev/null
  cmp "$G/demo.$f" "$G/demo.zero.$f"; rm "$G/demo.zero.$f"
done
md5sum "$G"/demo.* | sed "s#$G/##" > "$G/demo.md5"
rm "$G/demo.SAM_FULL" "$G/demo.DES"           # md5 only (SAM_FULL repeats the 2 MB of reads; DES == DES_FULL on this set)
gzip -9nf "$G/demo.SAM" "$G/demo.DES_FULL"
# second index: synthetic multi-strain reference (tools/gen_synth_ref.py 3 species x 4 strains x 300 kb, seed 7): unitigs with
# several reference positions, many secondary hits
SYN=$R/syn
mkdir -p "$SYN"
[ -s "$SYN/syn.fa" ] || python3 "$ROOT/tools/gen_synth_ref.py" "$SYN/syn.fa" 3 4 300000 7
[ -s "$SYN/idx/deSAMBA.bwt" ] || "$HERE/build_index.sh" "$SYN/syn.fa" "$SYN/idx"
[ -s "$R/sets/syn_long10.fq" ] || "$SIM" long "$SYN/syn.fa" 400 0.10 20261025 "$R/sets/syn_long10.fq"
[ -s "$R/sets/syn_short1.fq" ] || "$SIM" short "$SYN/syn.fa" 4000 0.01 20261026 "$R/sets/syn_short1.fq"
for s in syn_long10 syn_short1; do
  "$R/deSAMBA_zero" classify -t 1 -f DES_FULL "$SYN/idx" "$R/sets/$s.fq" -o "$G/$s.DES_FULL" 2>/dev/null; gzip -9nf "$G/$s.DES_FULL"
done
md5sum "$R"/sets/*.fq "$R/demo/ERR1050068.fastq" "$FA" "$SYN/syn.fa" | sed "s#$R/##" > "$G/inputs.md5"
ls -la "$G"
# l_ek = 17..20 (hash masks of 31..37 bits): the same synthetic index with its exist-k-mer tables re-written for a larger size
# class (oracle/rebuild_exk.c: the index builder's own rule, idx.c:998-1031; class 27 reproduces the builder's tables byte for
# byte), classified by the unmodified reference.  EK_CLASSES="28 30 32 34" (table bytes = 2^class each; 34 needs 32 GB of RAM + disk)
EKW=${EKW:-/tmp/dsb_ek}; mkdir -p "$EKW"
[ -x "$HERE/rebuild_exk" ] || gcc -O2 -w -o "$HERE/rebuild_exk" "$HERE/rebuild_exk.c"
"$HERE/rebuild_exk" "$SYN/idx" "$EKW/c27" 27 2>/dev/null && cmp "$EKW/c27/deSAMBA.exk0" "$SYN/idx/deSAMBA.exk0" && cmp "$EKW/c27/deSAMBA.exk1" "$SYN/idx/deSAMBA.exk1"
rm -rf "$EKW/c27"
for c in ${EK_CLASSES:-28 30 32}; do
  case $c in 28) ek=17;; 30) ek=18;; 32) ek=19;; 34) ek=20;; *) echo "class $c?"; exit 1;; esac
  "$HERE/rebuild_exk" "$SYN/idx" "$EKW/c$c" $c
  for s in syn_long10 syn_short1; do
    [ $c -ge 32 ] && [ $s = syn_long10 ] && continue
    "$R/deSAMBA_zero" classify -t 1 -f DES_FULL "$EKW/c$c" "$R/sets/$s.fq" -o "$G/$s.ek$ek.DES_FULL" 2>/dev/null; gzip -9nf "$G/$s.ek$ek.DES_FULL"
  done
  rm -rf "$EKW/c$c"
done
ls -la "$G"
