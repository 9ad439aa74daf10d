/*
 * sdt_readpack.c — multi-threaded FASTA/FASTQ parser + 2-bit packer (include/sdtpack.h).
 *
 * The files are mmap'ed.  Per call, records are first indexed serially (memchr from record to
 * record: tens of millions of records per second), then encoded and packed by a pool of threads,
 * each on a contiguous range of output records.  The parse rules restate the reference's
 * readseqInBuf / readseqfq (readseq1by1.c:122-178, 281-340); see sdtpack.h.
 */
#define _GNU_SOURCE
#include "../../include/sdtpack.h"

#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

typedef struct
{
	int fd;
	const char *map;
	size_t size, off;
	size_t *rec_off;	/* per indexed record: start of its header line */
	size_t *seq_off;	/* start of its sequence line */
	uint32_t *seq_len;	/* and that line's length in bytes (without the newline) */
	size_t n_idx, cap_idx;
	double rec_bytes;	/* running estimate of bytes per record (sizes the parallel index window) */
} file_t;

struct sdtpack_reader
{
	file_t f[2];
	int n_files, fastq, n_threads;
};

typedef struct
{
	const struct sdtpack_reader *r;
	/* output record j comes from file out_file(j), record out_rec(j) */
	uint64_t n_pair2;	/* first n_pair2 outputs alternate file 0 / file 1 */
	int tail_file;		/* the outputs after that all come from this file ... */
	uint64_t tail_first;	/* ... starting at this record index */
	uint64_t j0, j1;
	int max_read_len, n_kmer, reverse;
	uint8_t *packed, *nmask;
	uint32_t *lens;
	uint32_t stride;
	const uint8_t *lut;
} job_t;

static int file_open (file_t * f, const char *path)
{
	struct stat st;
	memset (f, 0, sizeof *f);
	f->fd = open (path, O_RDONLY);
	if (f->fd < 0)
		return -1;
	if (fstat (f->fd, &st) != 0)
		return -1;
	f->size = (size_t) st.st_size;
	if (f->size)
	{
		void *m = mmap (NULL, f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
		if (m == MAP_FAILED)
			return -1;
		f->map = (const char *) m;
		madvise (m, f->size, MADV_SEQUENTIAL);
	}
	return 0;
}

static void file_close (file_t * f)
{
	if (f->map)
		munmap ((void *) f->map, f->size);
	if (f->fd >= 0)
		close (f->fd);
	free (f->rec_off);
	free (f->seq_off);
	free (f->seq_len);
	memset (f, 0, sizeof *f);
	f->fd = -1;
}

static int idx_push (file_t * f, size_t rec, size_t off, size_t len)
{
	if (f->n_idx == f->cap_idx)
	{
		const size_t cap = f->cap_idx ? f->cap_idx * 2 : (1u << 14);
		size_t *r = (size_t *) realloc (f->rec_off, cap * sizeof (size_t));
		if (r)
			f->rec_off = r;
		size_t *a = r ? (size_t *) realloc (f->seq_off, cap * sizeof (size_t)) : NULL;
		if (a)
			f->seq_off = a;
		uint32_t *b = a ? (uint32_t *) realloc (f->seq_len, cap * sizeof (uint32_t)) : NULL;
		if (!b)
			return -1;
		f->seq_len = b;
		f->cap_idx = cap;
	}
	f->rec_off[f->n_idx] = rec;
	f->seq_off[f->n_idx] = off;
	f->seq_len[f->n_idx] = len > 0xFFFFFFFFu ? 0xFFFFFFFFu : (uint32_t) len;
	f->n_idx++;
	return 0;
}

/* end of the line that starts at p (position of its '\n', or size) */
static inline size_t line_end (const file_t * f, size_t p)
{
	const char *nl = p < f->size ? (const char *) memchr (f->map + p, '\n', f->size - p) : NULL;
	return nl ? (size_t) (nl - f->map) : f->size;
}

static inline size_t next_line (const file_t * f, size_t p)
{
	const size_t e = line_end (f, p);
	return e < f->size ? e + 1 : f->size;
}

/* first header line starting at or after p (p = start of a line); FASTQ headers are recognised by
 * "@..." with a '+' line two lines further down, so a quality line that begins with '@' is not one */
static size_t find_header (const file_t * f, int fastq, size_t p)
{
	while (p < f->size)
	{
		if (!fastq && f->map[p] == '>')
			return p;
		if (fastq && f->map[p] == '@')
		{
			const size_t plus = next_line (f, next_line (f, p));
			if (plus >= f->size || f->map[plus] == '+')
				return p;
		}
		p = next_line (f, p);
	}
	return f->size;
}

/* index the records whose header starts in [p, end) into `out` (a scratch file_t sharing the map);
 * p must be the start of a line; at most `want` records */
static int index_range (file_t * out, int fastq, size_t p, size_t end, uint64_t want)
{
	out->n_idx = 0;
	p = find_header (out, fastq, p);
	while (p < end && out->n_idx < want)
	{
		const size_t rec = p;
		size_t e;
		p = next_line (out, p);	/* header line */
		e = line_end (out, p);	/* first sequence line (may be empty at end of file) */
		if (idx_push (out, rec, p, e - p))
			return -1;
		p = e < out->size ? e + 1 : out->size;
		if (fastq)
			p = next_line (out, next_line (out, p));	/* '+' line and quality line */
		else
			p = find_header (out, 0, p);	/* further sequence lines of a FASTA record are ignored */
	}
	out->off = p;	/* start of the first record that was not indexed (or a position past `end`) */
	return 0;
}

typedef struct { file_t scratch; int fastq; size_t start, end; int rc; } idx_job_t;

static void *idx_worker (void *arg)
{
	idx_job_t *j = (idx_job_t *) arg;
	size_t p = j->start;
	if (p > 0 && j->scratch.map[p - 1] != '\n')
		p = next_line (&j->scratch, p);	/* chunk boundaries fall anywhere: move to the next line start */
	j->rc = p < j->end ? index_range (&j->scratch, j->fastq, p, j->end, ~0ull) : 0;
	if (p >= j->end)
		j->scratch.n_idx = 0;
	return NULL;
}

/* index up to `want` records from f->off on, in parallel over byte chunks; leaves f->off at the start
 * of the first unindexed record */
static int index_records (file_t * f, int fastq, uint64_t want, int n_threads)
{
	idx_job_t jobs[64];
	pthread_t th[64];
	size_t window, chunk;
	int t, nt, rc = 0;
	uint64_t taken = 0;

	f->n_idx = 0;
	if (f->off >= f->size || want == 0)
		return 0;
	if (f->rec_bytes <= 0)
	{	/* first call: measure a few records serially */
		file_t probe = *f;
		probe.rec_off = probe.seq_off = NULL; probe.seq_len = NULL; probe.cap_idx = probe.n_idx = 0;
		if (index_range (&probe, fastq, f->off, f->size, 64))
			return -1;
		f->rec_bytes = probe.n_idx ? (double) (probe.off - f->off) / (double) probe.n_idx : 256.0;
		free (probe.rec_off); free (probe.seq_off); free (probe.seq_len);
	}
	window = (size_t) ((double) want * f->rec_bytes * 1.02) + 4096;
	if (window > f->size - f->off)
		window = f->size - f->off;
	nt = n_threads;
	if (window < (1u << 20))
		nt = 1;
	chunk = (window + nt - 1) / nt;
	for (t = 0; t < nt; t++)
	{
		memset (&jobs[t], 0, sizeof jobs[t]);
		jobs[t].scratch.map = f->map;
		jobs[t].scratch.size = f->size;
		jobs[t].scratch.fd = -1;
		jobs[t].fastq = fastq;
		jobs[t].start = f->off + (size_t) t * chunk;
		jobs[t].end = f->off + ((size_t) (t + 1) * chunk < window ? (size_t) (t + 1) * chunk : window);
	}
	if (nt == 1)
		idx_worker (&jobs[0]);
	else
	{
		int started = 0;
		for (t = 0; t < nt; t++)
		{
			if (pthread_create (&th[t], NULL, idx_worker, &jobs[t]) != 0)
				break;
			started++;
		}
		for (t = started; t < nt; t++)
			idx_worker (&jobs[t]);
		for (t = 0; t < started; t++)
			pthread_join (th[t], NULL);
	}
	/* a FASTQ chunk may have synchronised inside the previous chunk's last record (its start is
	 * past that chunk's end); records are taken in order and duplicates by offset are dropped */
	for (t = 0; t < nt && !rc; t++)
	{
		file_t *s = &jobs[t].scratch;
		size_t i;
		if (jobs[t].rc)
			rc = -1;
		for (i = 0; i < s->n_idx && !rc && taken < want; i++)
		{
			if (f->n_idx && s->rec_off[i] <= f->rec_off[f->n_idx - 1])
				continue;
			if (idx_push (f, s->rec_off[i], s->seq_off[i], s->seq_len[i]))
				rc = -1;
			taken++;
		}
	}
	if (!rc)
	{
		if (f->n_idx == 0)
			f->off = f->off + window >= f->size ? f->size : f->off + window;	/* nothing but junk in the window */
		else
		{	/* continue right after the last record taken */
			const size_t last = f->n_idx - 1;
			size_t p = f->seq_off[last] + f->seq_len[last];
			p = p < f->size ? p + 1 : f->size;
			if (fastq)
				p = next_line (f, next_line (f, p));
			f->off = p;
			f->rec_bytes = 0.5 * f->rec_bytes + 0.5 * (double) (p - f->rec_off[0]) / (double) f->n_idx;
		}
	}
	for (t = 0; t < nt; t++)
	{
		free (jobs[t].scratch.rec_off); free (jobs[t].scratch.seq_off); free (jobs[t].scratch.seq_len);
	}
	return rc;
}

static void pack_one (const job_t * jb, const file_t * f, uint64_t rec, uint64_t j, char *codes)
{
	const char *s = f->map + f->seq_off[rec];
	uint32_t raw = f->seq_len[rec], n = 0, i;
	uint8_t *dst = jb->packed + j * jb->stride;
	if (raw > (uint32_t) jb->max_read_len)
		raw = (uint32_t) jb->max_read_len;	/* truncation counts raw characters (readseq1by1.c:147-148) */
	for (i = 0; i < raw; i++)
	{
		const uint8_t c = jb->lut[(unsigned char) s[i]];
		if (c != 0xFF)
			codes[n++] = (char) c;
	}
	if (jb->reverse && n)
	{	/* reverse2k: reverse, complement = x ^ 2 (an N stays an N here) */
		for (i = 0; i < n / 2; i++)
		{
			const char t = codes[i];
			codes[i] = codes[n - 1 - i];
			codes[n - 1 - i] = t;
		}
		for (i = 0; i < n; i++)
			if (codes[i] < 4)
				codes[i] ^= 2;
	}
	memset (dst, 0, jb->stride);
	for (i = 0; i < n; i++)
		dst[i >> 2] |= (uint8_t) ((codes[i] & 3) << (6 - 2 * (i & 3)));
	if (jb->nmask)
	{
		uint8_t *m = jb->nmask + j * (jb->stride / 2);
		memset (m, 0, jb->stride / 2);
		for (i = 0; i < n; i++)
			if (codes[i] == 4)
				m[i >> 3] |= (uint8_t) (0x80 >> (i & 7));
	}
	jb->lens[j] = n;
}

static void *worker (void *arg)
{
	const job_t *jb = (const job_t *) arg;
	char *codes = (char *) malloc ((size_t) jb->max_read_len + 8);
	uint64_t j;
	if (!codes)
		return (void *) 1;
	for (j = jb->j0; j < jb->j1; j++)
	{
		if (j < jb->n_pair2)
		{
			if (jb->r->n_files == 2)
				pack_one (jb, &jb->r->f[j & 1], j >> 1, j, codes);
			else
				pack_one (jb, &jb->r->f[0], j, j, codes);
		}
		else
			pack_one (jb, &jb->r->f[jb->tail_file], jb->tail_first + (j - jb->n_pair2), j, codes);
	}
	free (codes);
	return NULL;
}

int sdtpack_open (sdtpack_reader ** out, const char *path1, const char *path2, int fastq, int n_threads)
{
	struct sdtpack_reader *r;
	if (!out || !path1)
	{
		errno = EINVAL;
		return -1;
	}
	*out = NULL;
	r = (struct sdtpack_reader *) calloc (1, sizeof *r);
	if (!r)
		return -1;
	r->f[0].fd = r->f[1].fd = -1;
	r->n_files = path2 ? 2 : 1;
	r->fastq = fastq != 0;
	if (n_threads <= 0)
	{
		long c = sysconf (_SC_NPROCESSORS_ONLN);
		n_threads = c < 1 ? 1 : (c > 64 ? 64 : (int) c);
	}
	r->n_threads = n_threads;
	if (file_open (&r->f[0], path1) || (path2 && file_open (&r->f[1], path2)))
	{
		const int e = errno;
		sdtpack_close (r);
		errno = e;
		return -1;
	}
	*out = r;
	return 0;
}

void sdtpack_close (sdtpack_reader * r)
{
	if (!r)
		return;
	file_close (&r->f[0]);
	file_close (&r->f[1]);
	free (r);
}

int64_t sdtpack_next (sdtpack_reader * r, int max_read_len, int n_kmer, int reverse,
		      uint8_t * packed, uint32_t * lens, uint8_t * nmask, uint64_t max_reads, uint32_t stride_bytes)
{
	uint8_t lut[256];
	job_t jobs[64];
	pthread_t th[64];
	uint64_t n_out, per;
	int c, t, nt, failed = 0;
	job_t base;

	if (!r || !packed || !lens || max_read_len < 1 || (stride_bytes & 3) || (uint64_t) stride_bytes * 4 < (uint64_t) max_read_len ||
	    (n_kmer && !nmask) || max_reads < (uint64_t) r->n_files)
	{
		errno = EINVAL;
		return -1;
	}
	for (c = 0; c < 256; c++)
	{	/* readseq1by1.c:149-171 */
		if ((c == 'N' || c == 'n') && n_kmer)
			lut[c] = 4;
		else if (c >= 'a' && c <= 'z')
			lut[c] = (uint8_t) ((((c - 'a' + 'A') & 0x06) >> 1));
		else if (c >= 'A' && c <= 'Z')
			lut[c] = (uint8_t) ((c & 0x06) >> 1);
		else if (c == '.')
			lut[c] = 0;
		else
			lut[c] = 0xFF;
	}
	memset (&base, 0, sizeof base);
	base.r = r;
	base.max_read_len = max_read_len;
	base.n_kmer = n_kmer;
	base.reverse = reverse;
	base.packed = packed;
	base.lens = lens;
	base.nmask = n_kmer ? nmask : NULL;
	base.stride = stride_bytes;
	base.lut = lut;
	if (r->n_files == 1)
	{
		if (index_records (&r->f[0], r->fastq, max_reads, r->n_threads))
			return -1;
		base.n_pair2 = r->f[0].n_idx;
		n_out = base.n_pair2;
	}
	else
	{
		const uint64_t half = max_reads / 2;
		uint64_t n0, n1, np;
		if (index_records (&r->f[0], r->fastq, half, r->n_threads) || index_records (&r->f[1], r->fastq, half, r->n_threads))
			return -1;
		n0 = r->f[0].n_idx;
		n1 = r->f[1].n_idx;
		np = n0 < n1 ? n0 : n1;
		base.n_pair2 = 2 * np;
		n_out = 2 * np;
		if (n0 != n1)
		{	/* unequal windows: either one file ended (drain the other) or re-index the longer one shorter */
			const int longer = n0 > n1 ? 0 : 1, shorter = 1 - longer;
			if (r->f[shorter].off >= r->f[shorter].size)
			{
				base.tail_file = longer;
				base.tail_first = np;
				n_out += (longer == 0 ? n0 : n1) - np;
			}
			else
			{	/* the byte window of the shorter side held fewer records: give the surplus back */
				r->f[longer].off = r->f[longer].rec_off[np];
				r->f[longer].n_idx = np;
			}
		}
	}
	if (n_out == 0)
		return 0;
	nt = r->n_threads;
	if (n_out < 4096)
		nt = 1;
	if ((uint64_t) nt > n_out)
		nt = (int) n_out;
	per = (n_out + nt - 1) / nt;
	for (t = 0; t < nt; t++)
	{
		jobs[t] = base;
		jobs[t].j0 = (uint64_t) t * per;
		jobs[t].j1 = jobs[t].j0 + per < n_out ? jobs[t].j0 + per : n_out;
	}
	if (nt == 1)
		failed = worker (&jobs[0]) != NULL;
	else
	{
		int started = 0;
		for (t = 0; t < nt; t++)
		{
			if (pthread_create (&th[t], NULL, worker, &jobs[t]) != 0)
				break;
			started++;
		}
		for (t = started; t < nt; t++)	/* threads that could not be created: do their share here */
			failed |= worker (&jobs[t]) != NULL;
		for (t = 0; t < started; t++)
		{
			void *rv = NULL;
			pthread_join (th[t], &rv);
			failed |= rv != NULL;
		}
	}
	if (failed)
	{
		errno = ENOMEM;
		return -1;
	}
	return (int64_t) n_out;
}
