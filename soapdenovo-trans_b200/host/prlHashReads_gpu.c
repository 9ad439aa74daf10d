/*
 * prlHashReads_gpu.c — drop-in replacement for the reference's prlRead2HashTable
 * (src/prlHashReads.c:338-707, declared in src/inc/extfunc.h:156), the hashing stage of
 * `SOAPdenovo-Trans-{31,127}mer pregraph`.
 *
 * This file is compiled against the reference's own headers and linked into the reference binary
 * in place of the CPU hashing stage (host/Makefile renames the reference's symbol
 * to prlRead2HashTable_cpu with objcopy and links this definition instead).  Everything around it
 * is the unmodified reference: getopt in pregraph.c, the config parser scan_libInfo (lib.c:118),
 * file sequencing openNextFile (prlHashReads.c:808), and — after this function returns — the
 * consumers removeMinorOut / removeMinorTips / kmer2edges / prlRead2edge / output_vertex, which
 * walk the global KmerSets this function hands back in the reference's exact (set, slot) layout.
 *
 * What happens here, in the reference's order:
 *   1. same set-up and stdout lines as prlHashReads.c:355-367;
 *   2. every library file is read in the reference's order (openNextFile), records are parsed with
 *      the reference's rules (readseq1by1.c:122-178 FASTA, :281-340 FASTQ: first sequence line,
 *      base2int, '.' -> A, N -> G or 4 with -n, truncation to maxReadLen, reverse_seq), paired
 *      files are consumed alternately read1/read2 (prlHashReads.c:493-567);
 *   3. reads with len >= K+1 are 2-bit packed (seq.c:49-90 convention) into a pinned batch buffer;
 *      a full batch is one sdtgpu_push_reads call — the device replacement of
 *      `sendWorkSignal(2); sendWorkSignal(1);` (prlHashReads.c:468-469, 525-526, 562-563, 604-605,
 *      618-619);
 *   4. sdtgpu_finalize = deLowCov + Mark1in1outNode + the kmerFreq histogram (prlHashReads.c:689-699);
 *      this file writes <outfile>.kmerFreq in freqStat's format (:994-1023);
 *   5. sdtgpu_export_kmersets fills the global KmerSets.
 * Knobs that the reference CLI does not have come from the environment: SDTGPU_DEVICE,
 * SDTGPU_CAPACITY_HINT (expected distinct k-mers; 0 = grow on the device), SDTGPU_BATCH_READS.
 * BAM libraries (b=) are not supported by this path.
 */
#include "stdinc.h"
#include "newhash.h"
#include "extfunc.h"
#include "extvab.h"
#include "sdtgpu.h"

#ifdef MER127
#define SDT_KEY_WORDS 4
#endif
#ifdef MER63
#define SDT_KEY_WORDS 2
#endif
#ifdef MER31
#define SDT_KEY_WORDS 1
#endif

typedef struct
{
	sdtgpu_t *gpu;
	uint8_t *packed[2], *nmask[2];	/* pinned, double buffered */
	uint32_t *lens[2];
	int cur;
	uint32_t stride, mstride;
	uint64_t batch_reads, read_c, pushed_reads;
	long long instances;
	char *seq;	/* one parsed read, one base code per byte */
	char *line;
	size_t line_cap;
} hasher_t;

static void die (hasher_t * hs, const char *what, int rc)
{
	printf ("%s failed (%d): %s\n", what, rc, sdtgpu_last_error (hs ? hs->gpu : NULL));
	exit (-1);	/* the reference's error convention, check.c:31-34 */
}

static void flush_batch (hasher_t * hs)
{
	int rc;
	if (!hs->read_c)
		return;
	rc = sdtgpu_push_reads (hs->gpu, hs->packed[hs->cur], hs->lens[hs->cur], N_kmer ? hs->nmask[hs->cur] : NULL,
				hs->read_c, 0, hs->stride, hs->pushed_reads);
	if (rc)
		die (hs, "sdtgpu_push_reads", rc);
	hs->pushed_reads += hs->read_c;
	hs->read_c = 0;
	hs->cur ^= 1;
}

/* one read (len >= K+1) into the current batch */
static void add_read (hasher_t * hs, const char *seq, int len)
{
	uint8_t *dst = hs->packed[hs->cur] + hs->read_c * hs->stride;
	int i;
	memset (dst, 0, hs->stride);
	for (i = 0; i < len; i++)
		dst[i >> 2] |= (uint8_t) ((seq[i] & 3) << (6 - 2 * (i & 3)));
	if (N_kmer)
	{
		uint8_t *m = hs->nmask[hs->cur] + hs->read_c * hs->mstride;
		memset (m, 0, hs->mstride);
		for (i = 0; i < len; i++)
			if (seq[i] == 4)
				m[i >> 3] |= (uint8_t) (0x80 >> (i & 7));
	}
	hs->lens[hs->cur][hs->read_c] = (uint32_t) len;
	hs->instances += len - overlaplen + 1;
	if (++hs->read_c == hs->batch_reads)
		flush_batch (hs);
}

/* sequence line -> base codes, the rules of readseqInBuf/readseqfq (readseq1by1.c:147-171, 300-325) */
static int encode_line (const char *str, char *out)
{
	int n = 0, i, strL = (int) strlen (str);
	if (strL > maxReadLen)
		strL = maxReadLen;
	for (i = 0; i < strL; i++)
	{
		const char c = str[i];
		if ((c == 'N' || c == 'n') && N_kmer)
			out[n++] = 4;
		else if (c >= 'a' && c <= 'z')
			out[n++] = base2int (c - 'a' + 'A');
		else if (c >= 'A' && c <= 'Z')
			out[n++] = base2int (c);
		else if (c == '.')
			out[n++] = base2int ('A');
	}
	return n;
}

static int get_line (hasher_t * hs, FILE * fp)
{
	ssize_t got = getline (&hs->line, &hs->line_cap, fp);
	if (got < 0)
		return 0;
	while (got > 0 && (hs->line[got - 1] == '\n' || hs->line[got - 1] == '\r'))
		hs->line[--got] = '\0';
	return 1;
}

/* next record of a FASTA (fastq = 0) or FASTQ stream; returns 0 at end of file, else 1 and *len */
static int next_record (hasher_t * hs, FILE * fp, int fastq, int reverse, int *len)
{
	const char tag = fastq ? '@' : '>';
	int i;
	do
	{
		if (!get_line (hs, fp))
			return 0;
	}
	while (hs->line[0] != tag);
	if (!get_line (hs, fp))
		return 0;
	*len = encode_line (hs->line, hs->seq);
	if (fastq)
	{	/* '+' line and quality line */
		if (get_line (hs, fp))
			get_line (hs, fp);
	}
	if (reverse && *len)
	{	/* reverse2k, readseq1by1.c:749-764 */
		for (i = 0; i < *len / 2; i++)
		{
			char t = hs->seq[i];
			hs->seq[i] = hs->seq[*len - 1 - i];
			hs->seq[*len - 1 - i] = t;
		}
		for (i = 0; i < *len; i++)
			hs->seq[i] = int_comp (hs->seq[i]);
	}
	n_solexa++;
	return 1;
}

static void take_read (hasher_t * hs, int len, long long *i)
{
	if ((++(*i)) % 1000000 == 0)
		printf ("--- %lldth reads\n", *i);
	if (len < overlaplen + 1)	/* prlHashReads.c:507, 539, 592 */
		return;
	add_read (hs, hs->seq, len);
}

boolean prlRead2HashTable (char *libfile, char *outfile)
{
	hasher_t hs;
	long long i = 0;
	int libNo = 0, b, rc, len;
	time_t start_t, stop_t;
	const char *env;
	uint64_t hint = 0;
	int device = 0;
	int64_t freq[257];
	sdtgpu_stats st;
	char name[256];
	FILE *fo;
	sdtgpu_kmerset **sets;

	memset (&hs, 0, sizeof hs);
	WORDFILTER = createFilter (overlaplen);
	maxReadLen = 0;
	maxNameLen = 256;
	scan_libInfo (libfile);
	alloc_pe_mem (num_libs);
	if (!maxReadLen)
		maxReadLen = 100;
	maxReadLen4all = maxReadLen;
	printf ("In %s, %d libs, max seq len %d, max name len %d\n\n", libfile, num_libs, maxReadLen, maxNameLen);

	if ((env = getenv ("SDTGPU_DEVICE")))
		device = atoi (env);
	if ((env = getenv ("SDTGPU_CAPACITY_HINT")))
		hint = strtoull (env, NULL, 10);
	hs.batch_reads = 1u << 20;
	if ((env = getenv ("SDTGPU_BATCH_READS")) && atoll (env) > 0)
		hs.batch_reads = (uint64_t) atoll (env);
	rc = sdtgpu_create (&hs.gpu, device, overlaplen, SDT_KEY_WORDS, maxReadLen, hint, N_kmer ? SDTGPU_F_NKMER : 0);
	if (rc)
		die (NULL, "sdtgpu_create", rc);
	hs.stride = (uint32_t) (((maxReadLen + 3) / 4 + 3) / 4 * 4);
	hs.mstride = hs.stride / 2;
	for (b = 0; b < 2; b++)
	{
		if (sdtgpu_host_alloc ((void **) &hs.packed[b], hs.batch_reads * hs.stride) ||
		    sdtgpu_host_alloc ((void **) &hs.lens[b], hs.batch_reads * sizeof (uint32_t)) ||
		    sdtgpu_host_alloc ((void **) &hs.nmask[b], hs.batch_reads * hs.mstride))
			die (&hs, "sdtgpu_host_alloc", SDTGPU_ENOMEM);
	}
	hs.seq = (char *) ckalloc ((maxReadLen + 1) * sizeof (char));
	printf ("GPU pregraph hashing on device %d, K %d, %d-word keys\n", device, overlaplen, SDT_KEY_WORDS);

	time (&start_t);
	n_solexa = readNumBack = gradsCounter = 0;
	while (openNextFile (&libNo, 0, 1))
	{
		const int type = lib_array[libNo].curr_type, rev = lib_array[libNo].reverse;
		if (type == 4)
		{
			printf ("BAM input (b=) is not supported by the GPU pregraph path\n");
			exit (-1);
		}
		if (type == 1 || type == 2)
		{	/* paired files: read1, read2, read1, ... (prlHashReads.c:493-567) */
			int more1 = 1, more2 = 1;
			while (more1 || more2)
			{
				if (more1 && (more1 = next_record (&hs, lib_array[libNo].fp1, type == 2, rev, &len)))
					take_read (&hs, len, &i);
				if (more2 && (more2 = next_record (&hs, lib_array[libNo].fp2, type == 2, rev, &len)))
					take_read (&hs, len, &i);
			}
		}
		else
		{	/* p= / f= (FASTA) and q= (FASTQ): one stream (prlHashReads.c:574-611) */
			i = 0;	/* the reference restarts its progress counter here (:577) */
			while (next_record (&hs, lib_array[libNo].fp1, type == 6, rev, &len))
				take_read (&hs, len, &i);
		}
	}
	flush_batch (&hs);
	if ((rc = sdtgpu_sync (hs.gpu)))
		die (&hs, "sdtgpu_sync", rc);
	time (&stop_t);
	printf ("time spent on hash reads: %ds, %lld reads processed\n", (int) (stop_t - start_t), i);
	free_pe_mem ();
	free_libs ();

	time (&start_t);
	if ((rc = sdtgpu_finalize (hs.gpu, deLowKmer, freq, &st)))
		die (&hs, "sdtgpu_finalize", rc);
	printf ("%lli nodes allocated, %lli kmer in reads, %lli kmer processed\n", (long long) st.n_nodes, hs.instances, (long long) st.n_instances);
	if (deLowKmer)
		printf ("%lld kmer removed\n", (long long) st.n_removed);	/* deLowCov, :908 */
	printf ("%lld linear nodes\n", (long long) st.n_linear);	/* Mark1in1outNode, :991 */
	sprintf (name, "%s.kmerFreq", outfile);	/* freqStat, :994-1023 */
	fo = ckopen (name, "w");
	for (b = 1; b < 256; b++)
		fprintf (fo, "%lld\n", (long long) freq[b]);
	fclose (fo);
	time (&stop_t);
	printf ("time spent on marking linear nodes %ds\n", (int) (stop_t - start_t));

	time (&start_t);
	sets = (sdtgpu_kmerset **) ckalloc (thrd_num * sizeof (sdtgpu_kmerset *));
	if ((rc = sdtgpu_export_kmersets (hs.gpu, thrd_num, sets)))
		die (&hs, "sdtgpu_export_kmersets", rc);
	KmerSets = (KmerSet **) sets;	/* sdtgpu_kmerset is layout-identical to KmerSet (newhash.h:79-88) */
	time (&stop_t);
	printf ("time spent on handing the k-mer table back %ds\n", (int) (stop_t - start_t));
	fflush (stdout);

	for (b = 0; b < 2; b++)
	{
		sdtgpu_host_free (hs.packed[b]);
		sdtgpu_host_free (hs.lens[b]);
		sdtgpu_host_free (hs.nmask[b]);
	}
	free (hs.seq);
	free (hs.line);
	sdtgpu_destroy (hs.gpu);
	return 1;
}
