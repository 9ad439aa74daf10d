/*
 * ref_harness.c — TEST INFRASTRUCTURE (oracle/_ref), not product code.
 *
 * A small driver of our own that is linked against the UNMODIFIED reference objects
 * (every object of /root/reference/src/Makefile:6-14 except main.o) so that the reference's own
 * prlRead2HashTable (prlHashReads.c:338), hash_kmer (hashFunction.c:108), reverseComplement
 * (kmer.c:653) and put_kmerset (newhash.c:411) can be run and inspected.  It plays main.c's role
 * for the globals (global.h is included exactly once here, and kmerZero is defined here as
 * main.c:39-47 does).  Built only by oracle/Makefile, output only into oracle/_ref/.
 *
 * Sub-commands
 *   hash <cfg> <outprefix> -K k -p p -d d [-n] [-o dumpfile]
 *        run the reference's hashing stage; print one JSON line with wall seconds (clock_gettime
 *        around prlRead2HashTable) and node counts; optionally dump every occupied slot as a
 *        64-byte little-endian record:
 *          u32 set; u32 pad; u64 slot; u64 key[4] (most significant word first, zero padded);
 *          u32 l_links; u32 rword (r_links | linear<<24 | deleted<<25 | checked<<26 | single<<27
 *          | twin<<28 | inEdge<<30); u32 count; u32 pad
 *        preceded by a 32-byte header: "SDTDUMP1", u32 key_words, u32 n_sets, u64 n_records, u64 K
 *        and n_sets x {u64 size, u64 count, u64 max}.
 *   kat  reads lines "K SEQUENCE" from stdin, prints forward / reverse-complement words,
 *        KmerSmaller and hash_kmer of both (known-answer values for tests/golden).
 *   grow <n>  inserts n distinct keys into init_kmerset(1024,0.77f) and prints each (size,max).
 */
#include "stdinc.h"
#include "newhash.h"
#include "extfunc.h"
#include "global.h"
#include <time.h>

#ifdef MER127
Kmer kmerZero = { 0, 0, 0, 0 };
#define KEY_WORDS 4
#endif
#ifdef MER63
Kmer kmerZero = { 0, 0 };
#define KEY_WORDS 2
#endif
#ifdef MER31
Kmer kmerZero = 0;
#define KEY_WORDS 1
#endif

static void key_to_words (Kmer k, unsigned long long w[4])
{
	w[0] = w[1] = w[2] = w[3] = 0;
#ifdef MER127
	w[0] = k.high1; w[1] = k.low1; w[2] = k.high2; w[3] = k.low2;
#endif
#ifdef MER63
	w[2] = k.high; w[3] = k.low;
#endif
#ifdef MER31
	w[3] = k;
#endif
}

static Kmer key_from_words (const unsigned long long w[4])
{
	Kmer k;
#ifdef MER127
	k.high1 = w[0]; k.low1 = w[1]; k.high2 = w[2]; k.low2 = w[3];
#endif
#ifdef MER63
	k.high = w[2]; k.low = w[3];
#endif
#ifdef MER31
	k = w[3];
#endif
	return k;
}

static double now_s (void)
{
	struct timespec ts;
	clock_gettime (CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int cmd_hash (int argc, char **argv)
{
	char *cfg, *prefix, *dump = NULL;
	int i;
	if (argc < 2) { fprintf (stderr, "hash <cfg> <outprefix> ...\n"); return 2; }
	cfg = argv[0]; prefix = argv[1];
	overlaplen = 23; thrd_num = 8; deLowKmer = 0; N_kmer = 0;
	for (i = 2; i < argc; i++)
	{
		if (!strcmp (argv[i], "-K") && i + 1 < argc) overlaplen = atoi (argv[++i]);
		else if (!strcmp (argv[i], "-p") && i + 1 < argc) thrd_num = atoi (argv[++i]);
		else if (!strcmp (argv[i], "-d") && i + 1 < argc) deLowKmer = atoi (argv[++i]);
		else if (!strcmp (argv[i], "-n")) N_kmer = 1;
		else if (!strcmp (argv[i], "-o") && i + 1 < argc) dump = argv[++i];
		else { fprintf (stderr, "unknown option %s\n", argv[i]); return 2; }
	}
	/* same clamping as call_pregraph (pregraph.c:38-59) */
	if (overlaplen % 2 == 0) overlaplen++;
	if (overlaplen < 13) overlaplen = 13;
	if (overlaplen > 32 * KEY_WORDS - 1) overlaplen = 32 * KEY_WORDS - 1;

	double t0 = now_s ();
	prlRead2HashTable (cfg, prefix);
	double t1 = now_s ();

	unsigned long long n_nodes = 0, n_linear = 0, n_deleted = 0, n_single = 0, n_inst = 0;
	FILE *fo = NULL;
	if (dump)
	{
		fo = fopen (dump, "wb");
		if (!fo) { perror (dump); return 1; }
	}
	for (i = 0; i < thrd_num; i++) n_nodes += KmerSets[i]->count;
	if (fo)
	{
		unsigned int kw = KEY_WORDS, ns = thrd_num;
		unsigned long long nr = n_nodes, kk = overlaplen;
		fwrite ("SDTDUMP1", 1, 8, fo);
		fwrite (&kw, 4, 1, fo); fwrite (&ns, 4, 1, fo); fwrite (&nr, 8, 1, fo); fwrite (&kk, 8, 1, fo);
		for (i = 0; i < thrd_num; i++)
		{
			unsigned long long v[3] = { KmerSets[i]->size, KmerSets[i]->count, KmerSets[i]->max };
			fwrite (v, 8, 3, fo);
		}
	}
	for (i = 0; i < thrd_num; i++)
	{
		KmerSet *set = KmerSets[i];
		ubyte8 s;
		for (s = 0; s < set->size; s++)
		{
			if (is_kmer_entity_null (set->flags, s)) continue;
			kmer_t *n = set->array + s;
			n_linear += n->linear; n_deleted += n->deleted; n_single += n->single; n_inst += n->count;
			if (fo)
			{
				struct { unsigned int set, pad0; unsigned long long slot, key[4]; unsigned int l, r, c, pad1; } rec;
				memset (&rec, 0, sizeof rec);
				rec.set = i; rec.slot = s;
				key_to_words (n->seq, rec.key);
				rec.l = n->l_links;
				memcpy (&rec.r, (char *) &n->l_links + 4, 4);	/* the raw bit-field word (newhash.h:68-74) */
				rec.c = n->count;
				fwrite (&rec, sizeof rec, 1, fo);
			}
		}
	}
	if (fo) fclose (fo);
	printf ("REFJSON {\"seconds\": %.6f, \"nodes\": %llu, \"linear\": %llu, \"deleted\": %llu, \"single\": %llu, "
		"\"count_sum\": %llu, \"K\": %d, \"p\": %d, \"d\": %d, \"key_words\": %d}\n",
		t1 - t0, n_nodes, n_linear, n_deleted, n_single, n_inst, overlaplen, thrd_num, (int) deLowKmer, KEY_WORDS);
	return 0;
}

static int cmd_kat (void)
{
	char line[1024], seq[600];
	int K;
	while (fgets (line, sizeof line, stdin))
	{
		if (sscanf (line, "%d %599s", &K, seq) != 2) continue;
		unsigned long long w[4], r[4];
		int i;
		Kmer word = kmerZero;
		overlaplen = K;
		WORDFILTER = createFilter (K);
		for (i = 0; i < K; i++)
			word = KmerPlus (word, base2int (seq[i]));
		Kmer rc = reverseComplement (word, K);
		key_to_words (word, w); key_to_words (rc, r);
		printf ("%d %s %016llx %016llx %016llx %016llx %016llx %016llx %016llx %016llx %d %llu %llu\n", K, seq,
			w[0], w[1], w[2], w[3], r[0], r[1], r[2], r[3], (int) KmerSmaller (word, rc),
			(unsigned long long) hash_kmer (word), (unsigned long long) hash_kmer (rc));
	}
	return 0;
}

static int cmd_grow (int argc, char **argv)
{
	unsigned long long n = argc > 0 ? strtoull (argv[0], NULL, 10) : 100000, i, last = 0;
	unsigned long long x = 0x9E3779B97F4A7C15ULL;
	KmerSet *set = init_kmerset (1024, 0.77f);
	kmer_t *pos;
	printf ("%llu %llu\n", (unsigned long long) set->size, (unsigned long long) set->max);
	last = set->size;
	for (i = 0; i < n; i++)
	{
		unsigned long long w[4] = { 0, 0, 0, 0 };
		x = x * 6364136223846793005ULL + 1442695040888963407ULL;
		w[3] = (x >> 3) ^ i;
		put_kmerset (set, key_from_words (w), 4, 4, &pos);
		if (set->size != last)
		{
			last = set->size;
			printf ("%llu %llu\n", (unsigned long long) set->size, (unsigned long long) set->max);
		}
	}
	return 0;
}

int main (int argc, char **argv)
{
	if (argc < 2) { fprintf (stderr, "usage: %s hash|kat|grow ...\n", argv[0]); return 2; }
	if (!strcmp (argv[1], "hash")) return cmd_hash (argc - 2, argv + 2);
	if (!strcmp (argv[1], "kat")) return cmd_kat ();
	if (!strcmp (argv[1], "grow")) return cmd_grow (argc - 2, argv + 2);
	fprintf (stderr, "unknown sub-command %s\n", argv[1]);
	return 2;
}
