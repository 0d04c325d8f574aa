/*
 * rebuild_exk.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Re-writes the exist-k-mer tables of a deSAMBA index for a LARGER table size class, so that the l_ek = 17..20 / 31..37-bit
 * hash-mask paths of the classifier (set_ekmer_par, /root/reference/src/idx.c:966-982) can be exercised with a small
 * reference.  The reference's index builder picks the size class from the number of distinct 31-mers (get_EXIST_kmer,
 * idx.c:988-996: one eighth of a GiB below 2^31/9 k-mers, ..., 16 GiB at the top) -- an l_ek = 17 index needs a reference of
 * ~0.3 Gbp, l_ek = 20 tens of Gbp.  The tables themselves are a pure function of the unitig strings and the class
 * (idx.c:998-1031): bit hash64_1(kmer) & mask of table 0 and bit hash64_2(kmer) & mask of table 1, MSB-first, for every
 * l_ek-mer of every unitig.  This tool recovers the unitig text from the index's own BWT (inverse BWT by LF-mapping over the
 * 168-byte occ blocks, bwt.c:32-65, checked against the unitig lengths in .unv), applies that rule for the requested class
 * and writes <dst>/deSAMBA.{exki,exk0,exk1}; every other file of <dst> is a symlink into <src>.  The UNMODIFIED reference
 * (`deSAMBA classify`) then runs on <dst> exactly as on an index of that class: that is where the golden outputs of
 * tests/golden/*.ek17.* / *.ek18.* come from (oracle/make_golden.sh).
 *
 * usage: rebuild_exk <src_index_dir> <dst_index_dir> <log2 of the table size in bytes: 27 (the original class) .. 34>
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/stat.h>

static uint64_t hash64_1(uint64_t key)      /* utils.c:1067-1077 */
{
	key = (~key + (key << 21));
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8));
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4));
	key = key ^ key >> 28;
	key = (key + (key << 31));
	return key;
}
static uint64_t hash64_2(uint64_t key)      /* utils.c:1080-1091 */
{
	key += ~(key << 32);
	key ^= (key >> 22);
	key += ~(key << 13);
	key ^= (key >> 8);
	key += (key << 3);
	key ^= (key >> 15);
	key += ~(key << 27);
	key ^= (key >> 31);
	return key;
}

static void *slurp(const char *dir, const char *ext, size_t *n_out)
{
	char path[4096];
	snprintf(path, sizeof path, "%s/deSAMBA%s", dir, ext);
	FILE *f = fopen(path, "rb");
	if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(1); }
	fseek(f, 0, SEEK_END); size_t n = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
	void *p = malloc(n + 16);
	if (!p || fread(p, 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(1); }
	fclose(f);
	*n_out = n;
	return p;
}

/* occ block: u64 cnt[5] (A,C,G,T,# before the block) + 256 symbols as nibbles, low nibble first (bwt.c:32-41) */
static inline uint32_t bwt_sym(const uint8_t *blocks, uint64_t r)
{
	const uint8_t *b = blocks + (r >> 8) * 168 + 40;
	const uint32_t i = (uint32_t)(r & 255);
	return (b[i >> 1] >> ((i & 1) << 2)) & 0xf;
}
static inline uint64_t occ(const uint8_t *blocks, uint64_t r, uint32_t c)
{
	const uint8_t *blk = blocks + (r >> 8) * 168;
	uint64_t n = ((const uint64_t *)blk)[c];
	const uint8_t *b = blk + 40;
	const uint32_t in = (uint32_t)(r & 255);
	for (uint32_t i = 0; i < in; i++) n += (((b[i >> 1] >> ((i & 1) << 2)) & 0xf) == c);
	return n;
}

int main(int argc, char **argv)
{
	if (argc != 4) { fprintf(stderr, "usage: rebuild_exk <src_index_dir> <dst_index_dir> <log2 table bytes 27..34>\n"); return 2; }
	const char *src = argv[1], *dst = argv[2];
	const int lg = atoi(argv[3]);
	if (lg < 27 || lg > 34) { fprintf(stderr, "size class out of range\n"); return 2; }
	static const int l_ek_of[8] = {16, 17, 17, 18, 18, 19, 19, 20};      /* idx.c:970-977 */
	const int l_ek = l_ek_of[lg - 27];
	const uint64_t ek_size = 1ull << lg, mask = (1ull << (lg + 3)) - 1;

	size_t n_bwt, n_unv;
	uint8_t *bwt = (uint8_t *)slurp(src, ".bwt", &n_bwt);
	uint8_t *unv = (uint8_t *)slurp(src, ".unv", &n_unv);
	const uint64_t byteLen = *(uint64_t *)bwt;
	const uint8_t *blocks = bwt + 8;
	uint64_t rank[6];
	memcpy(rank, blocks + byteLen, 40);                                   /* A,C,G,T,# (bwt.c:80) */
	const uint64_t n_uni = *(uint64_t *)unv - 1;                          /* the file's last entry is the sentinel (idx.c:1123-1129) */
	const uint32_t *uni = (const uint32_t *)(unv + 8);                    /* {ref_list, length} */
	uint64_t text_len = 0;
	for (uint64_t u = 0; u < n_uni; u++) text_len += (uint64_t)uni[2 * u + 1] + 1;   /* unitig + '#' (the last one: '$') */
	fprintf(stderr, "[rebuild_exk] %llu unitigs, text of %llu symbols, class 2^%d bytes -> l_ek %d, mask %d bits\n",
	        (unsigned long long)n_uni, (unsigned long long)text_len, lg, l_ek, lg + 3);

	/* inverse BWT: rows 0 .. n_uni-1 are the suffixes that start with '#' / '$', '$' last (bwt.c:133-137); walking LF from the
	 * row of the suffix "$" yields the text right to left */
	uint8_t *text = (uint8_t *)malloc(text_len + 1);
	uint64_t r = n_uni - 1, pos = text_len - 1;
	text[pos] = 5;
	while (pos > 0) {
		const uint32_t c = bwt_sym(blocks, r);
		if (c > 4) { fprintf(stderr, "unexpected symbol %u at row %llu (pos %llu)\n", c, (unsigned long long)r, (unsigned long long)pos); return 1; }
		text[--pos] = (uint8_t)c;
		r = rank[c] + occ(blocks, r, c);
	}
	if (bwt_sym(blocks, r) != 5) { fprintf(stderr, "the walk did not end at the start of the text\n"); return 1; }
	{	/* check against the unitig table */
		uint64_t p = 0;
		for (uint64_t u = 0; u < n_uni; u++) {
			for (uint32_t k = 0; k < uni[2 * u + 1]; k++) if (text[p + k] > 3) { fprintf(stderr, "separator inside unitig %llu\n", (unsigned long long)u); return 1; }
			p += uni[2 * u + 1];
			if (text[p] != (u + 1 < n_uni ? 4 : 5)) { fprintf(stderr, "unitig %llu does not end where .unv says\n", (unsigned long long)u); return 1; }
			p++;
		}
	}

	/* the two bit tables (idx.c:998-1031) */
	uint8_t *t0 = (uint8_t *)calloc(ek_size, 1), *t1 = (uint8_t *)calloc(ek_size, 1);
	if (!t0 || !t1) { fprintf(stderr, "out of memory for 2 x %llu bytes\n", (unsigned long long)ek_size); return 1; }
	const uint64_t kmask = (l_ek == 32) ? ~0ull : ((1ull << (2 * l_ek)) - 1);
	uint64_t n_kmer = 0, p = 0;
	for (uint64_t u = 0; u < n_uni; u++) {
		const uint32_t len = uni[2 * u + 1];
		uint64_t kmer = 0;
		for (uint32_t i = 0; i < len; i++) {
			kmer = ((kmer << 2) | text[p + i]) & kmask;
			if (i + 1 >= (uint32_t)l_ek) {
				const uint64_t h1 = hash64_1(kmer) & mask, h2 = hash64_2(kmer) & mask;
				t0[h1 >> 3] |= (uint8_t)(0x80 >> (h1 & 7));
				t1[h2 >> 3] |= (uint8_t)(0x80 >> (h2 & 7));
				n_kmer++;
			}
		}
		p += (uint64_t)len + 1;
	}
	fprintf(stderr, "[rebuild_exk] %llu l_ek-mers hashed\n", (unsigned long long)n_kmer);

	mkdir(dst, 0755);
	static const char *keep[] = {".bwt", ".sa", ".acg", ".unv", ".ref_b", ".ref_i", ".ref_p"};
	char a[4096], b[4096], srcabs[4096];
	if (!realpath(src, srcabs)) { fprintf(stderr, "bad source directory\n"); return 1; }
	for (size_t k = 0; k < sizeof keep / sizeof *keep; k++) {
		snprintf(a, sizeof a, "%s/deSAMBA%s", srcabs, keep[k]); snprintf(b, sizeof b, "%s/deSAMBA%s", dst, keep[k]);
		unlink(b);
		if (access(a, R_OK) == 0 && symlink(a, b) != 0) { fprintf(stderr, "cannot link %s\n", b); return 1; }
	}
	snprintf(b, sizeof b, "%s/deSAMBA.exki", dst);
	FILE *f = fopen(b, "wb"); fwrite(&ek_size, 8, 1, f); fclose(f);
	snprintf(b, sizeof b, "%s/deSAMBA.exk0", dst);
	f = fopen(b, "wb"); if (fwrite(t0, 1, ek_size, f) != ek_size) { fprintf(stderr, "short write %s\n", b); return 1; } fclose(f);
	snprintf(b, sizeof b, "%s/deSAMBA.exk1", dst);
	f = fopen(b, "wb"); if (fwrite(t1, 1, ek_size, f) != ek_size) { fprintf(stderr, "short write %s\n", b); return 1; } fclose(f);
	return 0;
}
