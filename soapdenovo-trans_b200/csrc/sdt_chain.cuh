// sdt_chain.cuh — per-slice chains of record blocks: how super-k-mer records reach their slice.
//
// Round 1 wrote the records as one unordered stream, counted them per slice, scanned the counts and
// moved every record to its slice's run in a second kernel — 433 M scattered 32-byte stores that run
// at the B200's cold-line request rate (13.4 ms of a 95 ms step on C2) — and its experimental
// one-pass re-split of overflowing slices needed a count pass and a write pass to agree.
// Here a record goes to its slice the moment it is made: every slice (chain) owns a list of
// fixed-size blocks of CH_BLK records; appending is ONE 64-bit atomicAdd on the chain's head word
// (block << 32 | fill; the heads of all slices stay in L2) plus the record's own store.  The thread
// that finds a block exactly full links a fresh one from its CTA's private range of the block pool
// (no global traffic on that path either) and publishes it with an exchange; threads that arrive in
// between retry.  Nothing is counted in advance, so the same append serves the reads -> records
// kernel, the one-pass split of a slice that overflowed into sub-slices, and the receiving side of
// the multi-GPU exchange.  Afterwards the blocks of a chain are listed in order (scan of the blocks
// per chain + one pass over the blocks), and skm_merge_kernel streams every chain through shared
// memory, merges copies and leaves the survivors as one contiguous run per slice for the build.
#pragma once
#include "sdt_sliced.cuh"

namespace sdt {

static constexpr u32 CH_BLK = 32;		// records per block
static constexpr u32 CH_SB = 1024;		// blocks a CTA takes from the pool at a time
static constexpr u32 CH_LOW = 64;		// a CTA refills its range when fewer blocks than this are left
static constexpr u32 CH_NONE = 0xFFFFFFFFu;
static constexpr u64 OVF_RECORDS = 2, OVF_STORE = 4, OVF_FAILED = 8, OVF_FOREIGN = 16;	// Counters::overflow bits

struct SkmChains
{
	unsigned long long *head;	// [n_chains] block << 32 | records in it (> CH_BLK while a new block is being linked)
	u32 *bcount;			// [n_chains] blocks linked after the first (block c is chain c's first block)
	u32 *bchain, *bseq;		// [pool_blocks] chain of a linked block (CH_NONE: not linked), its number within the chain
	u32 *recs;			// the pool: pool_blocks x CH_BLK records
	unsigned long long *pool_cursor;	// next block nobody owns
	uint2 *cta_pool;		// [CTAs] what a CTA had left of its range when its last launch ended
	Counters *ctr;
	u64 pool_blocks;
	u32 n_chains, recw;
};

// a CTA's private range of blocks: s_pool[0] next, s_pool[1] end
__device__ __forceinline__ void pool_begin (const SkmChains &c, u32 *s_pool)
{
	if (threadIdx.x == 0)
	{
		const uint2 v = c.cta_pool[blockIdx.x];
		s_pool[0] = v.x;
		s_pool[1] = v.y;
	}
}
__device__ __forceinline__ void pool_end (const SkmChains &c, u32 *s_pool)
{
	if (threadIdx.x == 0)
		c.cta_pool[blockIdx.x] = make_uint2 (min (s_pool[0], s_pool[1]), s_pool[1]);
}
// thread 0, between barriers: keep at least `need` blocks in the range (what is left of the old range is dropped,
// so a refill takes several times `need`).  `need` is what a tile usually links, not a bound: pool_take has a slow path.
__device__ __forceinline__ void pool_refill (const SkmChains &c, u32 *s_pool, u32 need = CH_LOW)
{
	const u32 next = s_pool[0], end = s_pool[1];
	if (next + need <= end)
		return;
	const u32 take = max (CH_SB, 8 * need);
	const u64 base = atomicAdd (c.pool_cursor, (unsigned long long) take);
	if (base + take > c.pool_blocks)
	{	// the host enlarges the pool and emits again
		atomicOr (reinterpret_cast<unsigned long long *> (&c.ctr->overflow), (unsigned long long) OVF_RECORDS);
		s_pool[0] = s_pool[1] = 0;
		return;
	}
	s_pool[0] = (u32) base;
	s_pool[1] = (u32) base + take;
}

__device__ __forceinline__ u32 pool_take (const SkmChains &c, u32 *s_pool)
{
	const u32 i = atomicAdd (&s_pool[0], 1u);
	if (i < *reinterpret_cast<volatile u32 *> (&s_pool[1]))
		return i;
	// the CTA's range ran dry inside a tile (more blocks filled up at once than CH_LOW): straight from the pool
	const u64 b = atomicAdd (c.pool_cursor, 1ull);
	if (b < c.pool_blocks)
		return (u32) b;
	atomicOr (reinterpret_cast<unsigned long long *> (&c.ctr->overflow), (unsigned long long) OVF_RECORDS);
	return CH_NONE;
}

// where the next record of `chain` goes (nullptr: the pool is exhausted and the epoch is going to be emitted again).
// Split in two so that a caller can have the atomic of its next record in flight while it writes this one:
// chain_ticket draws a position, chain_place turns it into an address (drawing again if a block was being linked).
__device__ __forceinline__ u64 chain_ticket (const SkmChains &c, u32 chain)
{
	return atomicAdd (c.head + chain, 1ull);
}

// the thread that drew position CH_BLK links the chain's next block; the ticket becomes position 0 of that block
__device__ __forceinline__ bool chain_link (const SkmChains &c, u32 chain, u64 &ticket, u32 *s_pool)
{
	const u32 nb = pool_take (c, s_pool);
	if (nb == CH_NONE)
		return false;	// (the head stays full: whoever waits on it sees the overflow flag)
	c.bchain[nb] = chain;
	c.bseq[nb] = atomicAdd (c.bcount + chain, 1u);
	atomicExch (c.head + chain, ((u64) nb << 32) | 1ull);
	ticket = (u64) nb << 32;
	return true;
}

// `next` / `next_chain`: a ticket this thread has already drawn for its next record (or nullptr).  A thread that
// waits for somebody else's link must not sit on a link of its own, or two threads could wait for each other.
__device__ __forceinline__ u32 *chain_place (const SkmChains &c, u32 chain, u64 old, u32 *s_pool, u64 *next = nullptr, u32 next_chain = 0)
{
	for (u32 spin = 0;; spin++)
	{
		const u32 blk = (u32) (old >> 32), pos = (u32) old;
		if (pos < CH_BLK)
			return c.recs + ((u64) blk * CH_BLK + pos) * c.recw;
		if (pos == CH_BLK)
		{
			if (!chain_link (c, chain, old, s_pool))
				return nullptr;
			continue;
		}
		if (next && (u32) *next == CH_BLK)
			chain_link (c, next_chain, *next, s_pool);
		if ((spin & 15) == 15 && (*reinterpret_cast<volatile u64 *> (&c.ctr->overflow) & OVF_RECORDS))
			return nullptr;
		__nanosleep (32);
		old = atomicAdd (c.head + chain, 1ull);
	}
}

__device__ __forceinline__ u32 *chain_append (const SkmChains &c, u32 chain, u32 *s_pool)
{
	return chain_place (c, chain, chain_ticket (c, chain), s_pool);
}

// ---- chain state
__global__ void chain_init_kernel (unsigned long long *head, u32 *bcount, u32 n_chains, unsigned long long *pool_cursor)
{
	if (blockIdx.x == 0 && threadIdx.x == 0)
		*pool_cursor = n_chains;	// the pool behind the chains' first blocks
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_chains; i += gridDim.x * blockDim.x)
	{
		head[i] = (u64) i << 32;
		bcount[i] = 0;
	}
}

// blocks of every chain in order: blist[boff[chain] + number] = block (boff: exclusive scan of bcount)
__global__ void chain_list_kernel (const u32 *bchain, const u32 *bseq, const u64 *boff, u32 n_chains, const unsigned long long *pool_cursor, u64 pool_blocks, u32 *blist)
{
	const u64 n = min ((u64) *pool_cursor, pool_blocks);
	for (u64 b = n_chains + blockIdx.x * (u64) blockDim.x + threadIdx.x; b < n; b += (u64) gridDim.x * blockDim.x)
	{
		const u32 c = bchain[b];
		if (c != CH_NONE)
			blist[boff[c] + bseq[b]] = (u32) b;
	}
}

}	// namespace sdt
