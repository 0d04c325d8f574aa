#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle).  Builds the UNMODIFIED reference where it lies under
# /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun):
#   deSAMBA_stock  - reference sources, reference Makefile flags (-std=c99 -g -Wall -O3)   [O_stock, timed CPU baseline]
#   deSAMBA_zero   - same sources + -ftrivial-auto-var-init=zero                          [O_def, parity oracle; SURVEY 5.9/8c]
#   jellyfish      - Jellyfish 1.1.10 from the reference's zip, compiled by hand (no autotools here)
#   demo/          - viral-gs.fa, ERR1050068.fastq (unzipped demo data) and demo/idx (index built by the reference)
# No reference source is copied into the repo; only built binaries and data land in oracle/_ref/.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
[ -d "$REF/src" ] || { echo "no reference at $REF (GPU box?) - using prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$OUT/demo"
SRCS="$REF/src/lib/kthread.c $REF/src/lib/sam_format.c $REF/src/lib/utils.c $(ls $REF/src/*.c)"
if [ ! -x "$OUT/deSAMBA_stock" ]; then
  gcc -std=c99 -g -w -O3 -I "$REF/src/lib" -I "$REF/src" $SRCS -o "$OUT/deSAMBA_stock" -lm -lz -lpthread
fi
if [ ! -x "$OUT/deSAMBA_zero" ]; then
  gcc -std=c99 -g -w -O3 -ftrivial-auto-var-init=zero -I "$REF/src/lib" -I "$REF/src" $SRCS -o "$OUT/deSAMBA_zero" -lm -lz -lpthread
fi
if [ ! -x "$OUT/jellyfish" ]; then
  T=$(mktemp -d)
  python3 -c "import zipfile,sys; zipfile.ZipFile(sys.argv[1]).extractall(sys.argv[2])" "$REF/Jellyfish-1.1.10.zip" "$T"
  J=$(dirname "$(find "$T" -name Makefile.am | head -1)")
  printf '#define HAVE_INT128 1\n#define HAVE_EXECINFO_H 1\n#define HAVE_SYS_SYSCALL_H 1\n#define HAVE_SI_INT 1\n#define PACKAGE_STRING "jellyfish 1.1.10"\n#define PACKAGE_VERSION "1.1.10"\n#define PACKAGE_BUGREPORT "x"\n#define PACKAGE_NAME "jellyfish"\n#define VERSION "1.1.10"\n' > "$J/config.h"
  (cd "$J" && g++ -std=gnu++98 -O2 -w -fpermissive -DHAVE_CONFIG_H -D__STDC_CONSTANT_MACROS -D__STDC_FORMAT_MACROS -D__STDC_LIMIT_MACROS -I. \
    jellyfish/{yaggo.cpp,jellyfish.cc,stats_main.cc,hash_merge.cc,mer_counter.cc,histo_main.cc,dump_main.cc,query_main.cc,dump_fastq_main.cc,histo_fastq_main.cc,cite.cc,hash_fastq_merge.cc,square_binary_matrix.cc,err.cc,misc.cc,storage.cc,thread_exec.cc,time.cc,file_parser.cc,read_parser.cc,parse_read.cc,half.cpp,mapped_file.cc,parse_dna.cc,parse_quake.cc,parse_qual_dna.cc,sequence_parser.cc,seq_qual_parser.cc,backtrace.cc,floats.cc,dbg.cc,allocators_mmap.cc,dna_codes.cc} \
    -o "$OUT/jellyfish" -lpthread)
  rm -rf "$T"
fi
if [ ! -s "$OUT/demo/viral-gs.fa" ]; then
  python3 -c "import zipfile,sys; zipfile.ZipFile(sys.argv[1]).extractall(sys.argv[2])" "$REF/demo/viral-gs.zip" "$OUT/demo"
  python3 -c "import zipfile,sys; zipfile.ZipFile(sys.argv[1]).extractall(sys.argv[2])" "$REF/demo/ERR1050068.zip" "$OUT/demo"
fi
ls -la "$OUT" "$OUT/demo"
