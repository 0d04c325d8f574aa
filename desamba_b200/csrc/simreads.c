/*
 * simreads.c -- deterministic synthetic read generator for the benchmark / parity inputs.
 *
 *   simreads long  <ref.fa> <n_reads> <error_rate> <seed> <out.fq>   ONT-like: template length max(200, Gamma(k=2, theta=4000)),
 *                                                                     clipped to contig-2kb, 80 random bases prepended+appended,
 *                                                                     never within 1 kb of a contig end, contigs >= 2600 bp only
 *   simreads short <ref.fa> <n_reads> <error_rate> <seed> <out.fq>   150 bp, no padding, >= 100 bp from contig ends
 *   simreads mixed <ref.fa> <n_long> <n_short> <seed> <out.fq>       interleaved long (10 % / 30 % alternating) and short (1 %) reads
 * Errors are 1/3 substitution, 1/3 insertion, 1/3 deletion per erroneous base; strand 50/50; quality 'I'; FASTQ, 4 lines per read.
 * Contigs are sampled proportionally to their length.  PRNG: splitmix64-seeded xoshiro256**, so the output is identical everywhere.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <ctype.h>

static uint64_t s[4];
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rnd(void)
{
	uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
	s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
	return r;
}
static void seed_rng(uint64_t x)
{
	for (int i = 0; i < 4; i++) { x += 0x9e3779b97f4a7c15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; s[i] = z ^ (z >> 31); }
}
static inline double urand(void) { return ((rnd() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

typedef struct { char *seq; uint64_t len; } contig_t;
static contig_t *ctg; static size_t n_ctg;

static void load_fasta(const char *path)
{
	FILE *f = fopen(path, "r");
	if (!f) { fprintf(stderr, "simreads: cannot open %s\n", path); exit(1); }
	char *line = NULL; size_t cap = 0; ssize_t n; size_t m_ctg = 0; uint64_t m_seq = 0;
	while ((n = getline(&line, &cap, f)) > 0) {
		if (line[0] == '>') {
			if (n_ctg == m_ctg) { m_ctg = m_ctg ? m_ctg * 2 : 64; ctg = realloc(ctg, m_ctg * sizeof *ctg); }
			ctg[n_ctg].seq = NULL; ctg[n_ctg].len = 0; n_ctg++; m_seq = 0;
			continue;
		}
		if (!n_ctg) continue;
		contig_t *c = ctg + n_ctg - 1;
		while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) n--;
		if (c->len + n + 1 > m_seq) { m_seq = (c->len + n + 1) * 2; c->seq = realloc(c->seq, m_seq); }
		for (ssize_t i = 0; i < n; i++) c->seq[c->len + i] = toupper((unsigned char)line[i]);
		c->len += n;
	}
	free(line); fclose(f);
}

static const char ACGT[] = "ACGT";
static inline char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; } }

static char *buf; static size_t m_buf;
static inline void put(size_t *n, char c) { if (*n + 1 >= m_buf) { m_buf = m_buf ? m_buf * 2 : 1 << 16; buf = realloc(buf, m_buf); } buf[(*n)++] = c; }

/* cumulative-length table over eligible contigs */
static uint64_t *cum; static size_t *elig; static size_t n_elig;
static void build_table(uint64_t min_len)
{
	cum = malloc((n_ctg + 1) * 8); elig = malloc(n_ctg * sizeof *elig); n_elig = 0; uint64_t t = 0;
	for (size_t i = 0; i < n_ctg; i++) if (ctg[i].len >= min_len) { elig[n_elig] = i; t += ctg[i].len; cum[n_elig++] = t; }
	if (!n_elig) { fprintf(stderr, "simreads: no contig >= %lu bp\n", (unsigned long)min_len); exit(1); }
}
static contig_t *pick_contig(void)
{
	uint64_t x = (uint64_t)(urand() * cum[n_elig - 1]);
	size_t lo = 0, hi = n_elig - 1;
	while (lo < hi) { size_t mid = (lo + hi) / 2; if (cum[mid] > x) hi = mid; else lo = mid + 1; }
	return ctg + elig[lo];
}

static void emit(FILE *out, uint64_t id, contig_t *c, uint64_t pos, uint64_t tlen, double err, int pad)
{
	size_t n = 0;
	int rev = rnd() & 1;
	for (int i = 0; i < pad; i++) put(&n, ACGT[rnd() & 3]);
	for (uint64_t i = 0; i < tlen; i++) {
		char b = rev ? comp(c->seq[pos + tlen - 1 - i]) : c->seq[pos + i];
		if (urand() < err) {
			uint64_t k = rnd() % 3;
			if (k == 0) { char nb; do nb = ACGT[rnd() & 3]; while (nb == b); put(&n, nb); }
			else if (k == 1) { put(&n, ACGT[rnd() & 3]); put(&n, b); }
			/* k == 2: deletion */
		} else put(&n, b);
	}
	for (int i = 0; i < pad; i++) put(&n, ACGT[rnd() & 3]);
	buf[n] = 0;
	fprintf(out, "@r%lu\n%s\n+\n", (unsigned long)id, buf);
	memset(buf, 'I', n);
	fprintf(out, "%s\n", buf);
}

static void one_long(FILE *out, uint64_t id, double err)
{
	contig_t *c = pick_contig();
	double g = -4000.0 * (log(urand()) + log(urand()));
	uint64_t tlen = (uint64_t)(g < 200 ? 200 : g);
	if (tlen > c->len - 2000) tlen = c->len - 2000;
	uint64_t span = c->len - 2000 - tlen;               /* start in [1000, len-1000-tlen] */
	uint64_t pos = 1000 + (uint64_t)(urand() * (span + 1));
	emit(out, id, c, pos, tlen, err, 80);
}
static void one_short(FILE *out, uint64_t id, double err)
{
	contig_t *c = pick_contig();
	uint64_t span = c->len - 200 - 150;
	uint64_t pos = 100 + (uint64_t)(urand() * (span + 1));
	emit(out, id, c, pos, 150, err, 0);
}

int main(int argc, char **argv)
{
	if (argc < 7) { fprintf(stderr, "usage: simreads long|short <ref.fa> <n> <err> <seed> <out.fq>\n       simreads mixed <ref.fa> <n_long> <n_short> <seed> <out.fq>\n"); return 1; }
	load_fasta(argv[2]);
	FILE *out = fopen(argv[6], "w");
	if (!out) { fprintf(stderr, "simreads: cannot write %s\n", argv[6]); return 1; }
	setvbuf(out, NULL, _IOFBF, 1 << 22);
	seed_rng(strtoull(argv[5], NULL, 10));
	if (!strcmp(argv[1], "long")) {
		build_table(2600);
		uint64_t n = strtoull(argv[3], NULL, 10); double e = atof(argv[4]);
		for (uint64_t i = 0; i < n; i++) one_long(out, i, e);
	} else if (!strcmp(argv[1], "short")) {
		build_table(600);
		uint64_t n = strtoull(argv[3], NULL, 10); double e = atof(argv[4]);
		for (uint64_t i = 0; i < n; i++) one_short(out, i, e);
	} else if (!strcmp(argv[1], "mixed")) {
		build_table(2600);
		uint64_t nl = strtoull(argv[3], NULL, 10), ns = strtoull(argv[4], NULL, 10), il = 0, is = 0;
		while (il < nl || is < ns) {
			int take_long = (il < nl) && (is >= ns || urand() * (double)(nl - il + ns - is) < (double)(nl - il));
			if (take_long) { one_long(out, il + is, (il & 1) ? 0.30 : 0.10); il++; }
			else { one_short(out, il + is, 0.01); is++; }
		}
	} else { fprintf(stderr, "simreads: unknown mode %s\n", argv[1]); return 1; }
	fclose(out);
	return 0;
}
