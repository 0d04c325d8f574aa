/*
 * orc_classify_main.c -- TEST INFRASTRUCTURE ONLY.  CLI around the oracle restatement with the option surface of
 * `deSAMBA classify` (cly_mt.c:482-562), single-threaded (= the reference's -t 1 semantics incl. the running
 * max_read_l of cly.c:2958).  FASTQ only, plain text (4-line records or multi-line per kseq rules are not needed
 * for the generated inputs; '@name ...' / seq / '+' / qual).
 *   orc_classify [-l INT] [-s INT] [-r INT] [-f FMT] [-o FILE] [-c] <IndexDir> reads.fq...
 *   -c prints the algorithmic-byte counters to stderr.
 */
#include "desamba_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/time.h>

int main(int argc, char **argv)
{
	int l = 170, s = 64, r = 5, fmt = 1, c, counters = 0;
	FILE *out = stdout;
	while ((c = getopt(argc, argv, "t:l:r:f:o:s:c")) >= 0) {
		if (c == 'l') l = atoi(optarg);
		else if (c == 's') s = atoi(optarg);
		else if (c == 'r') r = atoi(optarg);
		else if (c == 'o') out = fopen(optarg, "w");
		else if (c == 'c') counters = 1;
		else if (c == 'f') {
			if (!strcmp(optarg, "SAM")) fmt = 1; else if (!strcmp(optarg, "SAM_FULL")) fmt = 2;
			else if (!strcmp(optarg, "DES")) fmt = 3; else if (!strcmp(optarg, "DES_FULL")) fmt = 4;
		}
	}
	if (optind + 2 > argc || !out) { fprintf(stderr, "usage: orc_classify [opts] <IndexDir> reads.fq...\n"); return 1; }
	orc_index ix;
	if (orc_index_load(&ix, argv[optind++])) return 2;
	orc_set_opts(&ix, l, s);
	orc_buff *buff = orc_buff_new();
	orc_result res; memset(&res, 0, sizeof res);
	size_t cap = 1 << 20; char *name = malloc(cap), *seq = malloc(cap), *plus = malloc(cap), *qual = malloc(cap);
	struct timeval t0, t1; gettimeofday(&t0, NULL);
	uint64_t n = 0;
	for (; optind < argc; optind++) {
		FILE *f = fopen(argv[optind], "r");
		if (!f) { fprintf(stderr, "cannot open %s\n", argv[optind]); return 3; }
		while (getline(&name, &cap, f) > 0) {
			if (getline(&seq, &cap, f) <= 0 || getline(&plus, &cap, f) <= 0 || getline(&qual, &cap, f) <= 0) break;
			name[strcspn(name, " \t\r\n")] = 0;
			seq[strcspn(seq, "\r\n")] = 0; qual[strcspn(qual, "\r\n")] = 0;
			uint32_t L = (uint32_t)strlen(seq);
			orc_classify_seq(&ix, seq, L, &res, buff);
			orc_write_result(out, &ix, &res, name + 1, seq, qual, L, fmt, r);
			n++;
		}
		fclose(f);
	}
	gettimeofday(&t1, NULL);
	double sec = (t1.tv_sec - t0.tv_sec) + (t1.tv_usec - t0.tv_usec) * 1e-6;
	fprintf(stderr, "%lu sequences processed in %.3fs (%.1f Kseq/m).\n", (unsigned long)n, sec, n / 1.0e3 / (sec / 60));
	if (counters)
		fprintf(stderr, "counters reads=%lu bases=%lu bit0=%lu bit1=%lu prefix=%lu occ=%lu locate=%lu getref=%lu getref_bytes=%lu hits=%lu\n",
		        (unsigned long)orc_cnt.n_reads, (unsigned long)orc_cnt.n_bases, (unsigned long)orc_cnt.n_bit0, (unsigned long)orc_cnt.n_bit1,
		        (unsigned long)orc_cnt.n_prefix, (unsigned long)orc_cnt.n_occ, (unsigned long)orc_cnt.n_locate, (unsigned long)orc_cnt.n_getref,
		        (unsigned long)orc_cnt.n_getref_bytes, (unsigned long)orc_cnt.n_hits);
	if (out != stdout) fclose(out);
	return 0;
}
