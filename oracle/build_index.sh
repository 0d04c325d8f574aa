#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle).  Build a deSAMBA index from a FASTA with the reference's own binaries
# in oracle/_ref (what /root/reference/build-index:73-110 does: jellyfish count -m 31 | kmersort | index).
# usage: oracle/build_index.sh <ref.fa> <index_dir>
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
R=$HERE/_ref
FA=$1; OUT=$2
[ -s "$OUT/deSAMBA.bwt" ] && { echo "index exists: $OUT" >&2; exit 0; }
mkdir -p "$OUT/jf"
SZ=$(stat -c %s "$FA"); HS=$((SZ*115/100))
"$R/jellyfish" count -m 31 -s $HS -t 8 -o "$OUT/jf/mer" "$FA"
if [ -e "$OUT/jf/mer_1" ]; then "$R/jellyfish" merge -o "$OUT/database.jdb" "$OUT"/jf/mer*; else mv "$OUT/jf/mer_0" "$OUT/database.jdb"; fi
rm -rf "$OUT/jf"
"$R/deSAMBA_stock" kmersort -k 31 -o "$OUT/kmer.srt" "$OUT/database.jdb" >&2
rm -f "$OUT/database.jdb"
"$R/deSAMBA_stock" index "$OUT/kmer.srt" "$FA" "$OUT" >&2
rm -f "$OUT/kmer.srt"
ls -la "$OUT" >&2
# the raw index directory is git- and gpurun-ignored (0.8 GB); the GPU box gets this archive (tests/oracle_binding.py unpacks it)
[ "$(basename "$OUT")" = idx ] && (cd "$(dirname "$OUT")" && tar cf - idx | gzip -1 > idx.tgz) || true
