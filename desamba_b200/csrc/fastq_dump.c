/* fastq_dump -- test tool for fastq_reader.h: prints the records of a FASTA/FASTQ file as  name \t sequence \t quality-length
 * usage: fastq_dump serial|parallel FILE [block_bytes] [threads] [margin_bytes]   (parallel falls back to the serial reader exactly
 * like the driver does; exit code 3 when it had to) */
#define _GNU_SOURCE
#include "fastq_reader.h"
#include <fcntl.h>
#include <sys/stat.h>

static int dump_serial(int fd, uint64_t from)
{
	stream_t st; memset(&st, 0, sizeof st); st.buf = malloc(SBUF); st.own_buf = st.buf; st.fd = fd; st.need_qual = 1;
	lseek(fd, (off_t)from, SEEK_SET);
	rec_t rec; memset(&rec, 0, sizeof rec);
	long L;
	while ((L = read_record(&st, &rec)) >= 0) {
		fwrite(rec.name, 1, rec.n_name, stdout); putchar('\t'); fwrite(rec.seq, 1, (size_t)L, stdout); printf("\t%zu\n", rec.n_qual);
	}
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 3) return 2;
	int fd = open(argv[2], O_RDONLY);
	if (fd < 0) return 2;
	if (!strcmp(argv[1], "serial")) return dump_serial(fd, 0);
	const int count_only = !strcmp(argv[1], "count");             /* timing aid: index only, print the totals */
	uint64_t n_rec = 0, n_base = 0;
	const uint64_t block = argc > 3 ? strtoull(argv[3], 0, 10) : (64u << 20);
	const int thr = argc > 4 ? atoi(argv[4]) : 4;
	const uint64_t margin = argc > 5 ? strtoull(argv[5], 0, 10) : (16u << 20);
	struct stat sb; fstat(fd, &sb);
	const uint64_t size = (uint64_t)sb.st_size;
	if (size == 0) return 0;
	char *buf = malloc(block + margin + 16);
	fq_list_t lists[64]; memset(lists, 0, sizeof lists);
	fq_rec_t *recs = NULL; size_t m_recs = 0;
	uint64_t pos = 0;
	char first = 0;
	if (pread(fd, &first, 1, 0) != 1 || first != '@') { dump_serial(fd, 0); return 3; }
	while (pos < size) {
		/* like the driver: block + margin bytes into the buffer, records that start inside the block are indexed */
		const uint64_t len = (size - pos < block + margin) ? size - pos : block + margin;
		if (fq_read_block(fd, pos, len, buf, thr)) return 2;
		const char *map = buf - pos;
		uint64_t next;
		const long n = fq_index_block(map, pos + len, pos + len == size, pos, pos + block, thr, lists, &recs, &m_recs, &next);
		if (n < 0) { fflush(stdout); dump_serial(fd, pos); return 3; }
		if (count_only) { for (long i = 0; i < n; i++) { n_rec++; n_base += recs[i].n_seq; } }
		else for (long i = 0; i < n; i++) {
			fwrite(map + recs[i].name, 1, recs[i].n_name, stdout); putchar('\t'); fwrite(map + recs[i].seq, 1, recs[i].n_seq, stdout); printf("\t%u\n", recs[i].n_seq);
		}
		pos = next;
	}
	if (count_only) printf("%llu records %llu bases\n", (unsigned long long)n_rec, (unsigned long long)n_base);
	return 0;
}
