/*
 * sdt_oracle.c — TEST INFRASTRUCTURE (see sdt_oracle.h).  Plain-C restatement of the
 * reference's pregraph k-mer hashing path, written from the reference's behaviour; each
 * function cites the reference file:line it follows (paths relative to /root/reference/src).
 * Parity is PINNED against the unmodified reference (tests/test_oracle.py, tests/golden/).
 */
#include "sdt_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ---------------------------------------------------------------- base codes (inc/def.h:39-42) */
int sdto_base2int (int ch)
{
	return (ch & 0x06) >> 1;	/* A=0 C=1 T=2 G=3 ; complement = x ^ 2 */
}

/* ---------------------------------------------------------------- 256-bit helpers (kmer.c:150-189) */
static sdto_kmer shl2 (sdto_kmer k)	/* KmerLeftBitMoveBy2, kmer.c:150 */
{
	k.w[0] = (k.w[0] << 2) | (k.w[1] >> 62);
	k.w[1] = (k.w[1] << 2) | (k.w[2] >> 62);
	k.w[2] = (k.w[2] << 2) | (k.w[3] >> 62);
	k.w[3] <<= 2;
	return k;
}

static sdto_kmer shr2 (sdto_kmer k)	/* KmerRightBitMoveBy2, kmer.c:171 */
{
	k.w[3] = (k.w[3] >> 2) | ((k.w[2] & 3) << 62);
	k.w[2] = (k.w[2] >> 2) | ((k.w[1] & 3) << 62);
	k.w[1] = (k.w[1] >> 2) | ((k.w[0] & 3) << 62);
	k.w[0] >>= 2;
	return k;
}

static sdto_kmer shr_bits (sdto_kmer k, int dis)	/* KmerRightBitMove, kmer.c:357: whole-value >> dis */
{
	int words = dis / 64, bits = dis % 64, i;
	sdto_kmer o = { {0, 0, 0, 0} };
	for (i = 3; i >= 0; i--)
	{
		int src = i - words;
		uint64_t v = 0;
		if (src >= 0)
		{
			v = k.w[src] >> bits;
			if (bits && src - 1 >= 0)
				v |= k.w[src - 1] << (64 - bits);
		}
		o.w[i] = v;
	}
	return o;
}

sdto_kmer sdto_create_filter (int K)	/* createFilter, kmer.c:313: the low 2K bits set */
{
	sdto_kmer f = { {0, 0, 0, 0} };
	int bits = 2 * K, i;
	for (i = 3; i >= 0 && bits > 0; i--, bits -= 64)
		f.w[i] = bits >= 64 ? ~0ULL : ((1ULL << bits) - 1);
	return f;
}

int sdto_kmer_smaller (sdto_kmer a, sdto_kmer b)	/* KmerSmaller, kmer.c:27: most significant word first */
{
	int i;
	for (i = 0; i < 3; i++)
		if (a.w[i] != b.w[i])
			return a.w[i] < b.w[i];
	return a.w[3] < b.w[3];
}

static int kmer_equal (sdto_kmer a, sdto_kmer b)	/* KmerEqual, kmer.c:111 */
{
	return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2] && a.w[3] == b.w[3];
}

sdto_kmer sdto_next_kmer (sdto_kmer prev, int ch, sdto_kmer filter)	/* nextKmer, kmer.c:209 */
{
	sdto_kmer w = shl2 (prev);
	int i;
	for (i = 0; i < 4; i++)
		w.w[i] &= filter.w[i];
	w.w[3] |= (uint64_t) (int64_t) (signed char) ch;
	return w;
}

sdto_kmer sdto_prev_kmer (sdto_kmer next, int ch, int K)	/* prevKmer, kmer.c:230 */
{
	sdto_kmer w = shr2 (next);
	int sh = 2 * (K - 1);
	w.w[3 - sh / 64] |= ((uint64_t) (unsigned char) ch) << (sh % 64);
	return w;
}

static uint64_t revcomp_word (uint64_t x)	/* one word of fastReverseComp, kmer.c:548-646 */
{
	x ^= 0xAAAAAAAAAAAAAAAAULL;
	x = ((x & 0x3333333333333333ULL) << 2) | ((x & 0xCCCCCCCCCCCCCCCCULL) >> 2);
	x = ((x & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((x & 0xF0F0F0F0F0F0F0F0ULL) >> 4);
	x = ((x & 0x00FF00FF00FF00FFULL) << 8) | ((x & 0xFF00FF00FF00FF00ULL) >> 8);
	x = ((x & 0x0000FFFF0000FFFFULL) << 16) | ((x & 0xFFFF0000FFFF0000ULL) >> 16);
	x = (x << 32) | (x >> 32);
	return x;
}

/* reverseComplement, kmer.c:653 -> fastReverseComp :548.  The reference complements and
 * group-reverses only the n = K/32+1 low words that can hold the k-mer, swaps them end for end,
 * then shifts the value right by 64n-2K bits; with K<32 the higher words are returned as they
 * came in (MER127/MER63 branch "if (seq_size < 32)").  key_words only bounds n. */
sdto_kmer sdto_reverse_complement (sdto_kmer word, int K, int key_words)
{
	int n = K / 32 + 1, i;
	sdto_kmer o = word;
	if (n > key_words)
		n = key_words;
	if (n == 1)
	{
		o.w[3] = revcomp_word (word.w[3]) >> (64 - 2 * K);
		return o;
	}
	for (i = 0; i < n; i++)
		o.w[3 - i] = revcomp_word (word.w[3 - (n - 1 - i)]);
	/* for n < 4 the untouched high words take part in the whole-value shift, as KmerRightBitMove does */
	return shr_bits (o, 64 * n - 2 * K);
}

/* ---------------------------------------------------------------- hash_kmer (hashFunction.c:83-122) */
static int32_t crc_tab[256];
static int crc_ready = 0;

static void crc_init (void)	/* the table at hashFunction.c:28-81 is the standard reflected CRC-32 table held in `int` */
{
	uint32_t n, c;
	int k;
	for (n = 0; n < 256; n++)
	{
		c = n;
		for (k = 0; k < 8; k++)
			c = (c & 1) ? (0xEDB88320U ^ (c >> 1)) : (c >> 1);
		crc_tab[n] = (int32_t) c;
	}
	crc_ready = 1;
}

uint64_t sdto_hash_kmer (sdto_kmer k, int key_words)
{
	unsigned char buf[32];
	int len = 8 * key_words, i;
	int32_t crc = 0;
	if (!crc_ready)
		crc_init ();
	/* raw bytes of the Kmer object: words in declaration order (high1, low1, high2, low2), each little endian */
	for (i = 0; i < key_words; i++)
		memcpy (buf + 8 * i, &k.w[4 - key_words + i], 8);
	crc = crc ^ (int32_t) 0xffffffff;
	for (i = 0; i < len; i++)
		crc = crc_tab[(crc ^ (int32_t) (signed char) buf[i]) & 0xff] ^ (crc >> 8);	/* `int` state: arithmetic shift (hashFunction.c:94) */
	crc = crc ^ (int32_t) 0xffffffff;
	return ((uint64_t) (int64_t) crc) & 0x0000000000ffffffULL;	/* KMER_HASH_MASK, hashFunction.c:25 */
}

/* ---------------------------------------------------------------- chopKmer4read (prlHashReads.c:164-310) */
int sdto_chop_read (const uint8_t *src, int len, int K, int key_words, int N_kmer,
		    sdto_kmer *kmers, uint8_t *prevc, uint8_t *nextc)
{
	const sdto_kmer zero = { {0, 0, 0, 0} };
	sdto_kmer filter = sdto_create_filter (K);
	sdto_kmer word = zero, bal;
	uint8_t *balseq;
	int index, j, bal_j, n_num = -1, out = 0;
	const uint8_t Invalid = 4;

	if (len < K + 1)	/* callers skip such reads (prlHashReads.c:507,539,592) */
		return 0;
	balseq = (uint8_t *) malloc ((size_t) len);
	for (index = 0; index < K; index++)	/* :179-197 */
	{
		if (key_words == 1)
		{
			word.w[3] <<= 2;
			word.w[3] += src[index];	/* MER31 uses += (:190-191) */
		}
		else
		{
			word = shl2 (word);
			word.w[3] |= src[index];	/* MER63/MER127 use |= (:182-187) */
		}
		if (src[index] == 4)
			n_num = K;
		else if (n_num > 0)
			n_num--;
	}
	for (j = 0; j < len; j++)	/* reverseComplementSeq, seq.c:93-110: int_comp = x ^ 2 */
		balseq[j] = (uint8_t) (src[len - 1 - j] ^ 0x02);
	bal = sdto_reverse_complement (word, K, key_words);	/* :201 */
	bal_j = len - 1 - K;	/* :203 */

	if (n_num > 0 && N_kmer)	/* :207-214 */
	{
		kmers[out] = zero; prevc[out] = Invalid; nextc[out] = Invalid; out++;
	}
	else if (sdto_kmer_smaller (word, bal))	/* :215-222 */
	{
		kmers[out] = word; prevc[out] = Invalid; nextc[out] = src[K]; out++;
	}
	else	/* :223-230 */
	{
		kmers[out] = bal; prevc[out] = balseq[bal_j]; nextc[out] = Invalid; out++;
	}
	for (j = 1; j <= len - K; j++)	/* :240-309 */
	{
		if (src[j - 1 + K] == 4)
			n_num = K;
		else if (n_num > 0)
			n_num--;
		else
			n_num = -1;
		if (n_num < K)
		{
			word = sdto_next_kmer (word, src[j - 1 + K], filter);
			bal_j = len - 1 - (j - 1 + K);
			bal = sdto_prev_kmer (bal, balseq[bal_j], K);
		}
		else
		{
			word = zero;
			bal = zero;
		}
		if (n_num > 0 && N_kmer)
		{
			kmers[out] = zero; prevc[out] = Invalid; nextc[out] = Invalid; out++;
		}
		else if (sdto_kmer_smaller (word, bal))
		{
			kmers[out] = word;
			prevc[out] = src[j - 1];
			nextc[out] = (j < len - K) ? src[j + K] : Invalid;
			out++;
		}
		else
		{
			kmers[out] = bal;
			prevc[out] = (bal_j > 0) ? balseq[bal_j - 1] : Invalid;
			nextc[out] = balseq[bal_j + K];
			out++;
		}
	}
	free (balseq);
	return out;
}

/* ---------------------------------------------------------------- KmerSet (newhash.c, inc/newhash.h) */
#define LINK_MAX 63u	/* MAX_KMER_COV, newhash.h:30 */
#define R_LINEAR  0x01000000u	/* bit-field word after l_links, newhash.h:68-74 */
#define R_DELETED 0x02000000u
#define R_SINGLE  0x08000000u
#define R_LINKS   0x00FFFFFFu

typedef struct { sdto_kmer seq; uint32_t l_links, rword, count; uint64_t first; } node_t;	/* first: test bookkeeping (ordinal of the first instance), not part of kmer_t */
typedef struct {
	node_t *array;
	uint32_t *flags;
	uint64_t size, count, max;
	double load_factor;
} set_t;

#define F_NULL(f, i)   (((f)[(i) >> 4] >> (((i) & 0x0f) << 1)) & 0x01)	/* newhash.h:47-53 */
#define F_EXISTS(f, i) (!(((f)[(i) >> 4] >> (((i) & 0x0f) << 1)) & 0x03))
#define F_SET_DEL(f, i)    ((f)[(i) >> 4] |= (0x02u << (((i) & 0x0f) << 1)))
#define F_CLEAR_NULL(f, i) ((f)[(i) >> 4] &= ~(0x01u << (((i) & 0x0f) << 1)))

static int is_prime_kh (uint64_t num)	/* newhash.c:116-141, including the float sqrt and `i < max` */
{
	uint64_t i, max;
	if (num < 4)
		return 1;
	if (num % 2 == 0)
		return 0;
	max = (uint64_t) sqrt ((float) num);
	for (i = 3; i < max; i += 2)
		if (num % i == 0)
			return 0;
	return 1;
}

uint64_t sdto_find_next_prime (uint64_t num)	/* newhash.c:143-158 */
{
	if (num % 2 == 0)
		num++;
	while (!is_prime_kh (num))
		num += 2;
	return num;
}

static set_t *set_init (uint64_t init_size, float load_factor)	/* init_kmerset, newhash.c:160-193 */
{
	set_t *s = (set_t *) malloc (sizeof (set_t));
	init_size = init_size < 3 ? 3 : sdto_find_next_prime (init_size);
	s->size = init_size;
	s->count = 0;
	s->max = s->size * load_factor;	/* float arithmetic, as in the reference (:176) */
	if (load_factor <= 0)
		load_factor = 0.25f;
	else if (load_factor >= 1)
		load_factor = 0.75f;
	s->load_factor = load_factor;
	s->array = (node_t *) calloc (s->size, sizeof (node_t));
	s->flags = (uint32_t *) malloc ((s->size + 15) / 16 * 4);
	memset (s->flags, 0x55, (s->size + 15) / 16 * 4);
	return s;
}

static uint64_t key_mod (const set_t *s, sdto_kmer k, int key_words)
{
	if (key_words == 1)	/* newhash.c:428 */
		return k.w[3] % s->size;
	if (key_words == 2)	/* newhash.c:423-425 */
	{
		unsigned __int128 t = ((unsigned __int128) k.w[2] << 64) | k.w[3];
		return (uint64_t) (t % s->size);
	}
	{	/* modular(), newhash.c:43-55: long division in 32-bit steps */
		uint64_t t;
		t = (k.w[0] % s->size) << 32 | (k.w[1] >> 32 & 0xffffffff);
		t = (t % s->size) << 32 | (k.w[1] & 0xffffffff);
		t = (t % s->size) << 32 | (k.w[2] >> 32 & 0xffffffff);
		t = (t % s->size) << 32 | (k.w[2] & 0xffffffff);
		t = (t % s->size) << 32 | (k.w[3] >> 32 & 0xffffffff);
		t = (t % s->size) << 32 | (k.w[3] & 0xffffffff);
		return t % s->size;
	}
}

static void set_encap (set_t *s, uint64_t num, int key_words)	/* encap_kmerset, newhash.c:293-409 */
{
	uint32_t *oldf, *newf;
	uint64_t i, n, old_size, hc;
	node_t key, tmp;

	if (s->count + num <= s->max)
		return;
	n = s->size;
	do
	{
		if (n < 0xFFFFFFFU)
			n <<= 1;
		else
			n += 0xFFFFFFU;
		n = sdto_find_next_prime (n);
	}
	while (n * s->load_factor < s->count + num);
	s->array = (node_t *) realloc (s->array, n * sizeof (node_t));
	if (!s->array)
	{
		fprintf (stderr, "-- Out of memory --\n");
		abort ();
	}
	newf = (uint32_t *) malloc ((n + 15) / 16 * 4);
	memset (newf, 0x55, (n + 15) / 16 * 4);
	old_size = s->size;
	s->size = n;
	s->max = n * s->load_factor;	/* double arithmetic here (:350) */
	oldf = s->flags;
	s->flags = newf;
	for (i = 0; i < old_size; i++)	/* in-place rehash with kick-out chains (:356-406) */
	{
		if (!F_EXISTS (oldf, i))
			continue;
		key = s->array[i];
		F_SET_DEL (oldf, i);
		while (1)
		{
			hc = key_mod (s, key.seq, key_words);
			while (!F_NULL (newf, hc))
			{
				hc++;
				if (hc == s->size)
					hc = 0;
			}
			F_CLEAR_NULL (newf, hc);
			if (hc < old_size && F_EXISTS (oldf, hc))
			{
				tmp = key;
				key = s->array[hc];
				s->array[hc] = tmp;
				F_SET_DEL (oldf, hc);
			}
			else
			{
				s->array[hc] = key;
				break;
			}
		}
	}
	free (oldf);
}

static void set_put (set_t *s, sdto_kmer seq, unsigned left, unsigned right, int key_words, uint64_t ordinal)	/* put_kmerset, newhash.c:411-462 */
{
	uint64_t hc;
	set_encap (s, 1, key_words);
	hc = key_mod (s, seq, key_words);
	while (1)
	{
		if (F_NULL (s->flags, hc))
		{	/* set_new_kmer, newhash.c:98-114 */
			node_t *m = s->array + hc;
			F_CLEAR_NULL (s->flags, hc);
			memset (m, 0, sizeof *m);
			m->seq = seq;
			m->rword = R_SINGLE;
			m->count = 1;
			m->first = ordinal;
			if (left < 4)
				m->l_links |= 1u << (6 * left);
			if (right < 4)
				m->rword |= 1u << (6 * right);
			s->count++;
			return;
		}
		if (kmer_equal (s->array[hc].seq, seq))
		{	/* update_kmer, newhash.c:71-96 */
			node_t *m = s->array + hc;
			uint32_t cov;
			m->count++;
			if (left < 4)
			{
				cov = (m->l_links >> (6 * left)) & 0x3f;
				if (cov < LINK_MAX)
					m->l_links += 1u << (6 * left);
			}
			if (right < 4)
			{
				cov = (m->rword >> (6 * right)) & 0x3f;
				if (cov < LINK_MAX)
					m->rword += 1u << (6 * right);
			}
			m->rword &= ~R_SINGLE;	/* :445 */
			return;
		}
		hc++;
		if (hc == s->size)
			hc = 0;
	}
}

/* ---------------------------------------------------------------- the hashing stage */
struct sdto_run {
	int K, key_words, thrd_num, N_kmer;
	set_t **sets;
	uint64_t instances, removed, linear, reads_seen;
	int max_read_len;
	int64_t freq[257];
	sdto_kmer *kbuf;
	uint8_t *pbuf, *nbuf;
	int bufcap;
};

sdto_run *sdto_run_create (int K, int key_words, int thrd_num, int N_kmer)
{
	sdto_run *r = (sdto_run *) calloc (1, sizeof *r);
	int i;
	r->K = K; r->key_words = key_words; r->thrd_num = thrd_num; r->N_kmer = N_kmer;
	r->max_read_len = 1 << 16;
	r->sets = (set_t **) malloc (sizeof (set_t *) * thrd_num);
	for (i = 0; i < thrd_num; i++)
		r->sets[i] = set_init (1024, 0.77f);	/* prlHashReads.c:402-416 */
	return r;
}

void sdto_run_push (sdto_run *r, const uint8_t *bases, const uint64_t *offsets, const uint32_t *lens, uint64_t n_reads)
{
	uint64_t t;
	int i, n;
	for (t = 0; t < n_reads; t++)
	{
		int len = (int) lens[t];
		if (len < r->K + 1)
			continue;
		if (len > r->bufcap)
		{
			r->bufcap = len + 64;
			r->kbuf = (sdto_kmer *) realloc (r->kbuf, sizeof (sdto_kmer) * r->bufcap);
			r->pbuf = (uint8_t *) realloc (r->pbuf, r->bufcap);
			r->nbuf = (uint8_t *) realloc (r->nbuf, r->bufcap);
		}
		n = sdto_chop_read (bases + offsets[t], len, r->K, r->key_words, r->N_kmer, r->kbuf, r->pbuf, r->nbuf);
		r->instances += n;
		for (i = 0; i < n; i++)
		{	/* owner = hashBan % thrd_num (prlHashReads.c:81); per-set order = instance order */
			uint64_t h = sdto_hash_kmer (r->kbuf[i], r->key_words);
			set_put (r->sets[h % r->thrd_num], r->kbuf[i], r->pbuf[i], r->nbuf[i], r->key_words,
				 (r->reads_seen + t) * (uint64_t) (r->max_read_len - r->K + 1) + (uint64_t) i);
		}
	}
	r->reads_seen += n_reads;
}

void sdto_run_finalize (sdto_run *r, int deLowKmer)
{
	int t, i;
	memset (r->freq, 0, sizeof r->freq);
	r->removed = r->linear = 0;
	for (t = 0; t < r->thrd_num; t++)
	{
		set_t *s = r->sets[t];
		uint64_t p;
		if (deLowKmer)	/* thread_delow, prlHashReads.c:844-887 (only run when deLowKmer != 0, :689) */
			for (p = 0; p < s->size; p++)
			{
				node_t *m;
				uint32_t rl;
				if (F_NULL (s->flags, p))
					continue;
				m = s->array + p;
				rl = m->rword & R_LINKS;
				for (i = 0; i < 4; i++)
				{
					uint32_t c = (m->l_links >> (6 * i)) & 0x3f;
					if (c > 0 && (int) c <= deLowKmer)
						m->l_links &= ~(0x3fu << (6 * i));
					c = (rl >> (6 * i)) & 0x3f;
					if (c > 0 && (int) c <= deLowKmer)
						rl &= ~(0x3fu << (6 * i));
				}
				m->rword = (m->rword & ~R_LINKS) | rl;
				if (m->l_links == 0 && rl == 0)
				{
					m->rword |= R_DELETED;
					r->removed++;
				}
			}
		for (p = 0; p < s->size; p++)	/* thread_mark, prlHashReads.c:911-967 */
		{
			node_t *m;
			int in_num = 0, out_num = 0, l_cvg = 0, r_cvg = 0;
			if (F_NULL (s->flags, p))
				continue;
			m = s->array + p;
			for (i = 0; i < 4; i++)
			{
				int c = (m->l_links >> (6 * i)) & 0x3f;
				if (c > 0) { in_num++; l_cvg += c; }
				c = (m->rword >> (6 * i)) & 0x3f;
				if (c > 0) { out_num++; r_cvg += c; }
			}
			if (m->rword & R_SINGLE)
				r->freq[1]++;
			else
				r->freq[l_cvg > r_cvg ? l_cvg : r_cvg]++;
			if (in_num == 1 && out_num == 1)
			{
				m->rword |= R_LINEAR;
				r->linear++;
			}
		}
	}
}

uint64_t sdto_run_nodes (const sdto_run *r)
{
	uint64_t n = 0;
	int t;
	for (t = 0; t < r->thrd_num; t++)
		n += r->sets[t]->count;
	return n;
}

uint64_t sdto_run_instances (const sdto_run *r) { return r->instances; }
uint64_t sdto_run_removed (const sdto_run *r) { return r->removed; }
uint64_t sdto_run_linear (const sdto_run *r) { return r->linear; }

void sdto_run_kmerfreq (const sdto_run *r, int64_t freq[257])
{
	memcpy (freq, r->freq, sizeof r->freq);
}

void sdto_run_set_info (const sdto_run *r, int set, uint64_t info[3])
{
	info[0] = r->sets[set]->size; info[1] = r->sets[set]->count; info[2] = r->sets[set]->max;
}

void sdto_run_dump (const sdto_run *r, sdto_record *out)
{
	uint64_t n = 0, p;
	int t;
	for (t = 0; t < r->thrd_num; t++)
	{
		const set_t *s = r->sets[t];
		for (p = 0; p < s->size; p++)
		{
			const node_t *m;
			if (F_NULL (s->flags, p))
				continue;
			m = s->array + p;
			memset (&out[n], 0, sizeof out[n]);
			out[n].set = (uint32_t) t;
			out[n].slot = p;
			memcpy (out[n].key, m->seq.w, 32);
			out[n].l_links = m->l_links;
			out[n].rword = m->rword;
			out[n].count = m->count;
			n++;
		}
	}
}

/* test bookkeeping: ordinal (read index * (max_read_len-K+1) + window) of each node's first
 * instance, in dump order; max_read_len must be set before the first push */
void sdto_run_set_max_read_len (sdto_run *r, int max_read_len) { r->max_read_len = max_read_len; }

void sdto_run_first_ordinals (const sdto_run *r, uint64_t *out)
{
	uint64_t n = 0, p;
	int t;
	for (t = 0; t < r->thrd_num; t++)
		for (p = 0; p < r->sets[t]->size; p++)
			if (!F_NULL (r->sets[t]->flags, p))
				out[n++] = r->sets[t]->array[p].first;
}

void sdto_run_destroy (sdto_run *r)
{
	int t;
	if (!r)
		return;
	for (t = 0; t < r->thrd_num; t++)
	{
		free (r->sets[t]->array);
		free (r->sets[t]->flags);
		free (r->sets[t]);
	}
	free (r->sets);
	free (r->kbuf); free (r->pbuf); free (r->nbuf);
	free (r);
}
