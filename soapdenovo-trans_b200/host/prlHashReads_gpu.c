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
 *   2. every library file is read in the reference's order (openNextFile); records are parsed and
 *      2-bit packed (seq.c:49-90 convention) by the multi-threaded reader of include/sdtpack.h, which
 *      restates the reference's rules (readseq1by1.c:122-178 FASTA, :281-340 FASTQ: first sequence
 *      line, base2int, '.' -> A, N -> G or 4 with -n, truncation to maxReadLen, reverse_seq; paired
 *      files alternate read1/read2, prlHashReads.c:493-567) straight into pinned batch buffers;
 *   3. a batch is one sdtgpu_push_reads call (reads shorter than K+1 are skipped on the device, as
 *      prlHashReads.c:507,539,592 skip them) — the device replacement of
 *      `sendWorkSignal(2); sendWorkSignal(1);` (prlHashReads.c:468-469, 525-526, 562-563, 604-605,
 *      618-619);
 *   4. sdtgpu_finalize = deLowCov + Mark1in1outNode + the kmerFreq histogram (prlHashReads.c:689-699);
 *      this file writes <outfile>.kmerFreq in freqStat's format (:994-1023);
 *   5. sdtgpu_export_kmersets fills the global KmerSets.
 * Knobs that the reference CLI does not have come from the environment: SDTGPU_DEVICES (comma
 * separated CUDA ordinals; with more than one, every GPU receives every batch and inserts the k-mers
 * it owns — the reference's own "each worker scans the batch and keeps hash % thrd_num == id",
 * prlHashReads.c:79-88 — and the nodes of all GPUs are merged at hand-back), SDTGPU_DEVICE (one
 * ordinal), SDTGPU_CAPACITY_HINT (expected distinct k-mers in total; optional), SDTGPU_DIRECT=1 (single-pass
 * insert into one open-addressing table instead of the sliced build), SDTGPU_BATCH_READS.
 * BAM libraries (b=) are not supported by this path.
 */
#include "stdinc.h"
#include "newhash.h"
#include "extfunc.h"
#include "extvab.h"
#include "sdtgpu.h"
#include "sdtpack.h"

#ifdef MER127
#define SDT_KEY_WORDS 4
#endif
#ifdef MER63
#define SDT_KEY_WORDS 2
#endif
#ifdef MER31
#define SDT_KEY_WORDS 1
#endif

#define SDT_MAX_GPUS 16

typedef struct
{
	sdtgpu_t *gpu;		/* gpus[0]: error reporting */
	sdtgpu_t *gpus[SDT_MAX_GPUS];
	int n_gpus;
	uint8_t *packed[2], *nmask[2];	/* pinned, double buffered */
	uint32_t *lens[2];
	int cur;
	uint32_t stride, mstride;
	uint64_t batch_reads, pushed_reads;
	long long instances;
} hasher_t;

static void die (hasher_t * hs, const char *what, int rc)
{
	printf ("%s failed (%d): %s\n", what, rc, sdtgpu_last_error (hs ? hs->gpu : NULL));
	exit (-1);	/* the reference's error convention, check.c:31-34 */
}

/* one file (or file pair) through the parser and the GPU; returns the number of reads seen */
static long long hash_file (hasher_t * hs, const char *path1, const char *path2, int fastq, int reverse, long long *progress)
{
	sdtpack_reader *rd;
	long long seen = 0;
	int64_t n;
	{	/* The reference as shipped reads a .gz library as PLAIN TEXT: the popen ("gzip -dc") branch of its
		 * openFile4read is commented out (readseq1by1.c:638-676).  This stage does the same bytes the same
		 * way; say so once, on stderr (stdout is compared with the reference's) */
		static int warned;
		const char *p[2] = { path1, path2 };
		int q;
		for (q = 0; q < 2 && !warned; q++)
		{
			size_t l = p[q] ? strlen (p[q]) : 0;
			while (l && p[q][l - 1] == ' ')
				l--;	/* the reference's trailing-space rule, readseq1by1.c:808-812 */
			if (l > 3 && !strncmp (p[q] + l - 3, ".gz", 3))
			{
				fprintf (stderr, "warning: %s looks gzip-compressed; like the reference, pregraph reads it as plain text — decompress it first\n", p[q]);
				warned = 1;
			}
		}
	}
	if (sdtpack_open (&rd, path1, path2, fastq, 0))
	{
		printf ("Cannot open %s%s%s\n", path1, path2 ? " / " : "", path2 ? path2 : "");
		exit (-1);
	}
	while ((n = sdtpack_next (rd, maxReadLen, N_kmer, reverse, hs->packed[hs->cur], hs->lens[hs->cur],
				  N_kmer ? hs->nmask[hs->cur] : NULL, hs->batch_reads, hs->stride)) > 0)
	{
		const uint32_t *lens = hs->lens[hs->cur];
		int64_t t;
		int rc;
		for (t = 0; t < n; t++)
			if ((int) lens[t] >= overlaplen + 1)	/* "kmer in reads", prlHashReads.c:516-518 */
				hs->instances += (int) lens[t] - overlaplen + 1;
		for (rc = 0, t = 0; t < hs->n_gpus && !rc; t++)	/* every GPU sees every batch and keeps what it owns */
			rc = sdtgpu_push_reads (hs->gpus[t], hs->packed[hs->cur], lens, N_kmer ? hs->nmask[hs->cur] : NULL,
						(uint64_t) n, 0, hs->stride, hs->pushed_reads);
		if (rc)
			die (hs, "sdtgpu_push_reads", rc);
		hs->pushed_reads += (uint64_t) n;
		hs->cur ^= 1;
		seen += n;
		n_solexa += n;
		while (*progress + 1000000 <= (long long) hs->pushed_reads)
		{
			*progress += 1000000;
			printf ("--- %lldth reads\n", *progress);
		}
	}
	if (n < 0)
	{
		printf ("reading %s failed\n", path1);
		exit (-1);
	}
	sdtpack_close (rd);
	return seen;
}

boolean prlRead2HashTable (char *libfile, char *outfile)
{
	hasher_t hs;
	long long i = 0, progress = 0;
	int libNo = 0, b, rc;
	time_t start_t, stop_t;
	const char *env;
	uint64_t hint = 0;
	uint32_t flags = N_kmer ? SDTGPU_F_NKMER : 0;
	int device = 0, devices[SDT_MAX_GPUS];
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
	devices[0] = device;
	hs.n_gpus = 1;
	if ((env = getenv ("SDTGPU_DEVICES")) && *env)
	{
		char *copy = strdup (env), *tok, *save = NULL;
		hs.n_gpus = 0;
		for (tok = strtok_r (copy, ",", &save); tok && hs.n_gpus < SDT_MAX_GPUS; tok = strtok_r (NULL, ",", &save))
			devices[hs.n_gpus++] = atoi (tok);
		free (copy);
		if (hs.n_gpus < 1)
		{
			devices[0] = device;
			hs.n_gpus = 1;
		}
	}
	/* the sliced build is the default: it needs no estimate of the distinct k-mers (SDTGPU_CAPACITY_HINT only tunes
	 * it); SDTGPU_DIRECT=1 selects the single-pass insert instead */
	if (!((env = getenv ("SDTGPU_DIRECT")) && atoi (env) > 0))
		flags |= SDTGPU_F_SLICED;
	for (b = 0; b < hs.n_gpus; b++)
	{
		rc = sdtgpu_create (&hs.gpus[b], devices[b], overlaplen, SDT_KEY_WORDS, maxReadLen,
				    hint ? hint / hs.n_gpus + hint / (8 * hs.n_gpus) + 1 : 0, flags);
		if (rc)
			die (NULL, "sdtgpu_create", rc);
		if (hs.n_gpus > 1 && (rc = sdtgpu_set_owner (hs.gpus[b], b, hs.n_gpus)))
			die (&hs, "sdtgpu_set_owner", rc);
	}
	hs.gpu = hs.gpus[0];
	hs.stride = (uint32_t) (((maxReadLen + 3) / 4 + 3) / 4 * 4);
	hs.mstride = hs.stride / 2;
	for (b = 0; b < 2; b++)
	{
		if (sdtgpu_host_alloc ((void **) &hs.packed[b], hs.batch_reads * hs.stride) ||
		    sdtgpu_host_alloc ((void **) &hs.lens[b], hs.batch_reads * sizeof (uint32_t)) ||
		    sdtgpu_host_alloc ((void **) &hs.nmask[b], hs.batch_reads * hs.mstride))
			die (&hs, "sdtgpu_host_alloc", SDTGPU_ENOMEM);
	}
	printf ("GPU pregraph hashing on %d device(s) (first: %d), K %d, %d-word keys\n", hs.n_gpus, devices[0], overlaplen, SDT_KEY_WORDS);

	time (&start_t);
	n_solexa = readNumBack = gradsCounter = 0;
	while (openNextFile (&libNo, 0, 1))
	{
		const LIB_INFO *lib = &lib_array[libNo];
		const int type = lib->curr_type, rev = lib->reverse, idx = lib->curr_index - 1;	/* openFileInLib advanced it */
		switch (type)
		{
		case 1:	i += hash_file (&hs, lib->a1_fname[idx], lib->a2_fname[idx], 0, rev, &progress); break;
		case 2:	i += hash_file (&hs, lib->q1_fname[idx], lib->q2_fname[idx], 1, rev, &progress); break;
		case 3:	i = hash_file (&hs, lib->p_fname[idx], NULL, 0, rev, &progress); break;	/* the reference restarts its counter for single files (:577) */
		case 5:	i = hash_file (&hs, lib->s_a_fname[idx], NULL, 0, rev, &progress); break;
		case 6:	i = hash_file (&hs, lib->s_q_fname[idx], NULL, 1, rev, &progress); break;
		default:
			printf ("BAM input (b=) is not supported by the GPU pregraph path\n");
			exit (-1);
		}
	}
	for (b = 0; b < hs.n_gpus; b++)
		if ((rc = sdtgpu_sync (hs.gpus[b])))
			die (&hs, "sdtgpu_sync", rc);
	time (&stop_t);
	printf ("time spent on hash reads: %ds, %lld reads processed\n", (int) (stop_t - start_t), i);
	free_pe_mem ();
	free_libs ();

	time (&start_t);
	memset (freq, 0, sizeof freq);
	memset (&st, 0, sizeof st);
	for (b = 0; b < hs.n_gpus; b++)
	{	/* per-GPU post-pass; counters and the histogram add up (owners are disjoint) */
		int64_t f1[257];
		sdtgpu_stats s1;
		int q;
		if ((rc = sdtgpu_finalize (hs.gpus[b], deLowKmer, f1, &s1)))
			die (&hs, "sdtgpu_finalize", rc);
		for (q = 0; q < 257; q++)
			freq[q] += f1[q];
		st.n_nodes += s1.n_nodes; st.n_instances += s1.n_instances; st.n_removed += s1.n_removed; st.n_linear += s1.n_linear;
	}
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
	if (hs.n_gpus == 1)
	{
		if ((rc = sdtgpu_export_kmersets (hs.gpu, thrd_num, sets)))
			die (&hs, "sdtgpu_export_kmersets", rc);
	}
	else
	{	/* merge the GPUs' nodes, then replay them into reference KmerSets on the host */
		sdtgpu_node *nodes = (sdtgpu_node *) malloc ((st.n_nodes ? st.n_nodes : 1) * sizeof (sdtgpu_node));
		uint64_t *last = (uint64_t *) calloc (thrd_num, sizeof (uint64_t)), got = 0, n1;
		if (!nodes || !last)
			die (&hs, "malloc for the merged hand-back", SDTGPU_ENOMEM);
		for (b = 0; b < hs.n_gpus; b++)
		{
			if ((rc = sdtgpu_export_nodes (hs.gpus[b], thrd_num, 0, nodes + got, st.n_nodes - got, &n1)))
				die (&hs, "sdtgpu_export_nodes", rc);
			got += n1;
		}
		rc = sdtgpu_last_ordinals (hs.gpu, thrd_num, last);	/* every GPU saw every batch: any one of them knows */
		if (rc && rc != SDTGPU_ESTATE)
			die (&hs, "sdtgpu_last_ordinals", rc);
		if ((rc = sdtgpu_build_kmersets (nodes, got, SDT_KEY_WORDS, thrd_num, rc == SDTGPU_ESTATE ? NULL : last, sets)))
			die (&hs, "sdtgpu_build_kmersets", rc);
		free (nodes);
		free (last);
	}
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
	for (b = 0; b < hs.n_gpus; b++)
		sdtgpu_destroy (hs.gpus[b]);
	return 1;
}
