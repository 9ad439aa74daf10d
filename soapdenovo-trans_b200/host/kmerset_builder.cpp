// kmerset_builder.cpp — host half of the hand-back: turns exported nodes into reference KmerSets.
//
// The consumers (removeMinorOut cutTipPreGraph.c:1012, kmer2edges node2edge.c:46, prlRead2edge
// prlRead2path.c:817) sweep the sets in (set, slot) order and their results depend on that order
// (SURVEY.md §7.3-1), so the hand-back must place every node in exactly the slot the reference's
// own init_kmerset(1024,0.77f) + put_kmerset sequence would have produced (newhash.c:160-193,
// 293-462).  That layout is a pure function of, per set, the order in which DISTINCT keys first
// appear — which the GPU table records as the node's minimum instance ordinal.  This file replays
// that order through a model of the reference table's growth policy and in-place re-hash, tracking
// node indices instead of 24/48-byte records, and then writes the kmer_t arrays.
#include "../../include/sdtgpu.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

typedef uint64_t u64;

bool looks_prime (u64 num)
{	// is_prime_kh, newhash.c:116-141: trial division that stops below (ubyte8) sqrt ((float) num)
	if (num < 4)
		return true;
	if (num % 2 == 0)
		return false;
	const u64 lim = (u64) std::sqrt ((double) (float) num);
	for (u64 i = 3; i < lim; i += 2)
		if (num % i == 0)
			return false;
	return true;
}

u64 next_size (u64 num)
{	// find_next_prime_kh, newhash.c:143-158
	if (num % 2 == 0)
		num++;
	while (!looks_prime (num))
		num += 2;
	return num;
}

u64 key_mod (const u64 key[4], int key_words, u64 size)
{
	if (key_words == 1)
		return key[3] % size;	// newhash.c:428
	if (key_words == 2)
	{	// newhash.c:423-425
		const unsigned __int128 v = ((unsigned __int128) key[2] << 64) | key[3];
		return (u64) (v % size);
	}
	u64 t = key[0] % size;	// modular(), newhash.c:43-55: base-2^32 long division
	for (int i = 1; i < 4; i++)
	{
		t = ((t << 32) | (key[i] >> 32)) % size;
		t = ((t << 32) | (key[i] & 0xffffffffull)) % size;
	}
	return t;
}

// Model of ONE reference KmerSet being filled in arrival order.  Slots hold the RANK of a node in
// that order (or -1); the keys are kept in a compact rank-ordered array so that the replay touches
// 8-32 bytes per node instead of a 56-byte export record, and both loops prefetch the random slot
// (and key) they are about to need — the replay is bound by cache misses, not arithmetic.
struct SetModel
{
	std::vector<int32_t> slot;	// rank or -1   (a set holds far fewer than 2^31 nodes: ranks fit 32 bits)
	std::vector<u64> keys;		// key_words words per rank, most significant first (only the words of the build)
	u64 size = 0, count = 0, max = 0;
	double load = 0;
	int key_words = 1;

	u64 mod_of (size_t rank) const
	{
		u64 k4[4] = { 0, 0, 0, 0 };
		for (int w = 0; w < key_words; w++)
			k4[4 - key_words + w] = keys[rank * key_words + w];
		return key_mod (k4, key_words, size);
	}

	void init ()
	{
		const float lf = 0.77f;	// prlHashReads.c:409-415
		size = next_size (1024);
		max = (u64) ((float) size * lf);	// float product, newhash.c:176
		load = lf;
		count = 0;
		slot.assign (size, -1);
	}

	void grow (u64 num)
	{	// encap_kmerset, newhash.c:293-409
		if (count + num <= max)
			return;
		u64 n = size;
		do
		{
			if (n < 0xFFFFFFFull)
				n <<= 1;
			else
				n += 0xFFFFFFull;
			n = next_size (n);
		}
		while ((double) n * load < (double) (count + num));
		const u64 old_size = size;
		slot.resize (n, -1);
		// state of the OLD positions: 1 = holds a node not yet re-placed, 0 = free for the taking
		std::vector<uint8_t> pending (old_size);
		for (u64 i = 0; i < old_size; i++)
			pending[i] = slot[i] >= 0;
		std::vector<uint8_t> taken (n, 0);	// new-table occupancy ("null" bit cleared)
		size = n;
		max = (u64) ((double) n * load);
		const u64 AHEAD = 12;
		for (u64 i = 0; i < old_size; i++)
		{
			if (i + 2 * AHEAD < old_size && slot[i + 2 * AHEAD] >= 0)
				__builtin_prefetch (&keys[(size_t) slot[i + 2 * AHEAD] * key_words]);
			if (i + AHEAD < old_size && pending[i + AHEAD])
			{
				const u64 h = mod_of ((size_t) slot[i + AHEAD]);
				__builtin_prefetch (&taken[h]);
				__builtin_prefetch (&slot[h]);
			}
			if (!pending[i])
				continue;
			int32_t carry = slot[i];
			pending[i] = 0;
			for (;;)
			{
				u64 hc = mod_of ((size_t) carry);
				while (taken[hc])
					if (++hc == size)
						hc = 0;
				taken[hc] = 1;
				if (hc < old_size && pending[hc])
				{	// the target still holds an un-moved node: swap it out and keep going
					std::swap (carry, slot[hc]);
					pending[hc] = 0;
				}
				else
				{
					slot[hc] = carry;
					break;
				}
			}
		}
		// positions of the old range that were vacated and not re-taken are empty now
		for (u64 i = 0; i < old_size; i++)
			if (!taken[i])
				slot[i] = -1;
	}

	// put_kmerset for keys known to be absent (newhash.c:411-440), ranks first..last-1 in order
	void put_all (size_t n_ranks)
	{
		const size_t AHEAD = 12;
		for (size_t r = 0; r < n_ranks; r++)
		{
			grow (1);
			if (r + AHEAD < n_ranks)
				__builtin_prefetch (&slot[mod_of (r + AHEAD)]);	// (a growth in between just wastes the hint)
			u64 hc = mod_of (r);
			while (slot[hc] >= 0)
				if (++hc == size)
					hc = 0;
			slot[hc] = (int32_t) r;
			count++;
		}
	}
};

size_t node_bytes (int key_words) { return key_words == 1 ? 24 : (key_words == 2 ? 32 : 48); }

}	// namespace

extern "C" int sdtgpu_build_kmersets (const sdtgpu_node *nodes, uint64_t n_nodes, int key_words, int thrd_num,
				      const uint64_t *set_last_ordinal, sdtgpu_kmerset **sets)
{
	if ((!nodes && n_nodes) || !sets || thrd_num < 1 || (key_words != 1 && key_words != 2 && key_words != 4))
		return SDTGPU_EINVAL;
	std::vector<std::vector<int64_t>> order (thrd_num);
	for (u64 i = 0; i < n_nodes; i++)
	{
		if (nodes[i].set >= (uint32_t) thrd_num)
			return SDTGPU_EINVAL;
		order[nodes[i].set].push_back ((int64_t) i);
	}
	std::vector<int> status (thrd_num, SDTGPU_OK);
	auto work = [&](int t) {
		std::vector<int64_t> &ord = order[t];
		if (ord.size () >= 0x7FFFFFFFull)
		{
			status[t] = SDTGPU_ERANGE;
			sets[t] = nullptr;
			return;
		}
		{	// arrival order = ascending first-instance ordinal; sort (ordinal, index) pairs, not indices
			std::vector<std::pair<u64, int64_t>> keyed (ord.size ());
			for (size_t i = 0; i < ord.size (); i++)
				keyed[i] = std::make_pair ((u64) nodes[ord[i]].ordinal, ord[i]);
			std::sort (keyed.begin (), keyed.end ());
			for (size_t i = 0; i < ord.size (); i++)
				ord[i] = keyed[i].second;
		}
		SetModel m;
		m.key_words = key_words;
		m.keys.resize (ord.size () * (size_t) key_words);
		for (size_t i = 0; i < ord.size (); i++)
			for (int w = 0; w < key_words; w++)
				m.keys[i * key_words + w] = nodes[ord[i]].key[4 - key_words + w];
		m.init ();
		m.put_all (ord.size ());
		// put_kmerset evaluates encap_kmerset on EVERY call, also for instances of known keys
		// (newhash.c:415): if instances of this set arrived after its last new key, one more
		// growth step may have happened.
		if (set_last_ordinal && !ord.empty () && set_last_ordinal[t] > nodes[ord.back ()].ordinal)
			m.grow (1);
		sdtgpu_kmerset *s = (sdtgpu_kmerset *) malloc (sizeof (sdtgpu_kmerset));
		const size_t nb = node_bytes (key_words);
		char *array = (char *) calloc (m.size, nb);
		const size_t fwords = (m.size + 15) / 16;
		uint32_t *flags = (uint32_t *) malloc (fwords * 4);
		if (!s || !array || !flags)
		{
			free (s); free (array); free (flags);
			sets[t] = nullptr;
			status[t] = SDTGPU_ENOMEM;
			return;
		}
		memset (flags, 0x55, fwords * 4);	// every entry "null" (newhash.c:190-191, newhash.h:47)
		for (u64 p = 0; p < m.size; p++)
		{
			if (m.slot[p] < 0)
				continue;
			const sdtgpu_node &n = nodes[ord[(size_t) m.slot[p]]];
			char *rec = array + p * nb;
			memcpy (rec, &n.key[4 - key_words], 8 * (size_t) key_words);	// Kmer words in declaration order (def.h:45-59)
			memcpy (rec + 8 * key_words, &n.l_links, 4);
			memcpy (rec + 8 * key_words + 4, &n.rword, 4);
			memcpy (rec + 8 * key_words + 8, &n.count, 4);
			flags[p >> 4] &= ~(1u << ((p & 15) << 1));	// clear_kmer_entity_null
		}
		s->array = array;
		s->flags = flags;
		s->size = m.size;
		s->count = m.count;
		s->max = m.max;
		s->load_factor = m.load;
		s->iter_ptr = 0;
		sets[t] = s;
	};
	std::vector<std::thread> pool;
	for (int t = 0; t < thrd_num; t++)
		pool.emplace_back (work, t);
	for (auto &th : pool)
		th.join ();
	for (int t = 0; t < thrd_num; t++)
		if (status[t])
		{
			sdtgpu_free_kmersets (sets, thrd_num);
			return status[t];
		}
	return SDTGPU_OK;
}

extern "C" void sdtgpu_free_kmersets (sdtgpu_kmerset **sets, int thrd_num)
{
	if (!sets)
		return;
	for (int t = 0; t < thrd_num; t++)
		if (sets[t])
		{
			free (sets[t]->array);
			free (sets[t]->flags);
			free (sets[t]);
			sets[t] = nullptr;
		}
}
