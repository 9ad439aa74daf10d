// sdt_device.cuh — device-side building blocks of the pregraph k-mer hashing path (sm_100a).
//
// What the reference does per window (prlHashReads.c:164-310 chopKmer4read, kmer.c:209-265
// nextKmer/prevKmer, kmer.c:548-653 reverseComplement, kmer.c:27 KmerSmaller, newhash.c:411-462
// put_kmerset, newhash.c:71-114 update_kmer/set_new_kmer) is re-designed here for the GPU:
//   * no rolling state: each thread extracts its window straight from the 2-bit packed read tile
//     in shared memory with funnel shifts, and gets the reverse complement with brev + pair swap;
//   * the table is one open-addressing array of 32-byte (K<=63) or 64-byte (K<=127) slots, one
//     DRAM sector per slot for K<=63, claimed with a 64-bit CAS (K<=31), a 128-bit CAS (K<=63) or
//     a 128-bit lock word + publish (K<=127); all payload updates are commutative atomics, so
//     insertion order is free and the result is bit-identical to the reference.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sdt {

typedef unsigned long long u64;
typedef unsigned int u32;

static constexpr u64 EMPTY64 = ~0ull;
static constexpr u64 LOCKED64 = ~0ull - 1;	// word 1 of a locked 4-word key's first half
static constexpr u64 ORD40_NONE = (1ull << 40) - 1;
static constexpr u32 LINK_SAT = 63;	// MAX_KMER_COV, inc/newhash.h:30

// ------------------------------------------------------------------------------------------------
// memory primitives (all table traffic bypasses L1: the table is mutated by L2 atomics).
// .L2::64B: by default a missing load makes the B200 L2 fetch the whole 128-byte line (127 B of DRAM
// reads per random 16-byte load, tools/ldvar_bench.cu); the 64-byte prefetch size is the smallest
// the ISA offers and halves that (cudaLimitMaxL2FetchGranularity has no effect on this part).
__device__ __forceinline__ void ld256 (const void *p, u64 &a, u64 &b, u64 &c, u64 &d)
{
	asm volatile ("ld.global.relaxed.gpu.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void ld128 (const void *p, u64 &a, u64 &b)
{
	asm volatile ("ld.global.relaxed.gpu.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st256 (void *p, u64 a, u64 b, u64 c, u64 d)
{
	asm volatile ("st.global.relaxed.gpu.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void st128 (void *p, u64 a, u64 b)
{
	asm volatile ("st.global.relaxed.gpu.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}
// 128-bit compare-and-swap (ATOMG.E.CAS.128 on sm_90+); returns the old value in (oa, ob)
__device__ __forceinline__ bool cas128 (void *p, u64 ca, u64 cb, u64 na, u64 nb, u64 &oa, u64 &ob)
{
	asm volatile ("{\n\t.reg .b128 c, n, o;\n\tmov.b128 c, {%2,%3};\n\tmov.b128 n, {%4,%5};\n\t"
		      "atom.global.relaxed.gpu.cas.b128 o, [%6], c, n;\n\tmov.b128 {%0,%1}, o;\n\t}"
		      : "=l"(oa), "=l"(ob) : "l"(ca), "l"(cb), "l"(na), "l"(nb), "l"(p) : "memory");
	return oa == ca && ob == cb;
}
__device__ __forceinline__ void red_add_u32 (u32 *p, u32 v)
{
	asm volatile ("red.global.relaxed.gpu.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_min_u64 (u64 *p, u64 v)
{
	asm volatile ("red.global.relaxed.gpu.min.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u64 fmix64 (u64 k)
{
	k ^= k >> 33; k *= 0xff51afd7ed558ccdull;
	k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull;
	k ^= k >> 33;
	return k;
}

// ------------------------------------------------------------------------------------------------
// keys: w[0] most significant 64-bit word; the k-mer sits in the low 2K bits (kmer.c convention)
template <int W> struct Key { u64 w[W]; };

template <int W> __device__ __forceinline__ bool key_less (const Key<W> &a, const Key<W> &b)
{	// KmerSmaller, kmer.c:27
#pragma unroll
	for (int i = 0; i < W - 1; i++)
		if (a.w[i] != b.w[i])
			return a.w[i] < b.w[i];
	return a.w[W - 1] < b.w[W - 1];
}

template <int W> __device__ __forceinline__ u64 key_hash (const Key<W> &k)
{
	u64 h = fmix64 (k.w[W - 1]);
#pragma unroll
	for (int i = W - 2; i >= 0; i--)
		h = fmix64 (h ^ (k.w[i] + 0x9E3779B97F4A7C15ull));
	return h;
}

__device__ __forceinline__ u64 slot_of (u64 h, u64 cap) { return __umul64hi (h, cap); }
// owner rank for the multi-GPU exchange: decorrelated from slot_of (which uses the high bits)
__device__ __forceinline__ u32 owner_of (u64 h, u32 n_ranks) { return __umulhi ((u32) (h * 0x9E3779B97F4A7C15ull >> 32), n_ranks); }

// Forward k-mer of the window that ENDS (exclusive) at base `end` of a read whose packed bits are
// the big-endian 32-bit words rd[...] (word q holds bases 16q..16q+15, first base in the top bits).
// Words below index 0 may be read (their bits are masked away); the tile is padded for that.
template <int W> __device__ __forceinline__ void extract_fwd (const u32 *rd, int end, int K, Key<W> &k)
{
	const int e = 2 * end, qe = e >> 5, r = e & 31;
	u32 o[2 * W];
#pragma unroll
	for (int i = 0; i < 2 * W; i++)
	{
		const int bits = 2 * K - 32 * i;	// bits of the k-mer that fall in output word i
		u32 v = 0;
		if (bits > 0)
		{
			v = __funnelshift_l (rd[qe - i], rd[qe - i - 1], r);
			if (bits < 32)
				v &= (1u << bits) - 1u;
		}
		o[i] = v;
	}
#pragma unroll
	for (int m = 0; m < W; m++)
		k.w[W - 1 - m] = ((u64) o[2 * m + 1] << 32) | o[2 * m];
}

__device__ __forceinline__ u64 revcomp64 (u64 x)
{	// complement (A0 C1 T2 G3: x ^ 2 per base, inc/def.h:42) and reverse the 32 two-bit groups
	x = __brevll (x ^ 0xAAAAAAAAAAAAAAAAull);
	return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// reverse complement of a right-aligned K-mer held in W words (reverseComplement, kmer.c:653)
template <int W> __device__ __forceinline__ void revcomp (const Key<W> &f, int K, Key<W> &rc)
{
	u64 t[W];	// left-aligned reverse complement, t[0] most significant
#pragma unroll
	for (int i = 0; i < W; i++)
		t[i] = revcomp64 (f.w[W - 1 - i]);
	const int s = 64 * W - 2 * K;	// right shift that re-aligns it; 2 <= s, and s < 64 unless W == 4
	if constexpr (W == 1)
		rc.w[0] = t[0] >> s;
	else if constexpr (W == 2)
	{
		rc.w[0] = t[0] >> s;
		rc.w[1] = (t[1] >> s) | (t[0] << (64 - s));
	}
	else
	{
		if (s >= 64)
		{	// K <= 96: one whole word drops out (b = s - 64 in 0..62)
			const int b = s - 64;
			rc.w[0] = 0;
			rc.w[1] = t[0] >> b;
			rc.w[2] = b ? ((t[1] >> b) | (t[0] << (64 - b))) : t[1];
			rc.w[3] = b ? ((t[2] >> b) | (t[1] << (64 - b))) : t[2];
		}
		else
		{
			rc.w[0] = t[0] >> s;
			rc.w[1] = (t[1] >> s) | (t[0] << (64 - s));
			rc.w[2] = (t[2] >> s) | (t[1] << (64 - s));
			rc.w[3] = (t[3] >> s) | (t[2] << (64 - s));
		}
	}
}

__device__ __forceinline__ u32 base_at (const u32 *rd, int b)
{
	return (rd[b >> 4] >> (30 - 2 * (b & 15))) & 3u;
}

// ------------------------------------------------------------------------------------------------
// table slots: key words + one 16-byte payload, everything of a slot in one 32-byte DRAM sector
// for K <= 63 (two sectors for K <= 127).
//   payload word 0 = ord40 << 24 | L24      L24 = the four left  link counters, 6 bits each
//   payload word 1 = count32 << 32 | R24    R24 = the four right link counters
// Link counters are the reference's saturating 6-bit counters (update_kmer, newhash.c:71-96),
// count its wrapping u32 (newhash.c:75); ord is the smallest instance ordinal seen, kept only so
// that the export can reproduce the reference's slot order.
struct __align__(16) Payload { u64 ordL; u32 R; u32 count; };
struct __align__(32) Slot1 { u64 key; u64 pad; Payload p; };
struct __align__(32) Slot2 { u64 key[2]; Payload p; };
struct __align__(64) Slot4 { u64 key[4]; Payload p; u64 pad[2]; };

template <int W> struct SlotOf;
template <> struct SlotOf<1> { typedef Slot1 type; };
template <> struct SlotOf<2> { typedef Slot2 type; };
template <> struct SlotOf<4> { typedef Slot4 type; };

static constexpr u64 PAYLOAD0_INIT = ORD40_NONE << 24;

__device__ __forceinline__ u32 sat_inc24 (u32 links, u32 base)
{
	if (base < 4 && ((links >> (6 * base)) & 63) < LINK_SAT)
		links += 1u << (6 * base);
	return links;
}

// One instance into a slot's payload.  Fast path: ONE 128-bit CAS carries count+1, both link
// increments and the ordinal minimum (a hash-table insert costs what its L2 requests cost:
// tools/randacc_bench.cu, ld + CAS128 17 G/s vs ld + CAS64 + RED 12 G/s on random slots).
// If the CAS loses a race the key is contended (a hot k-mer): fall back to per-field atomics,
// which never retry on account of count (RED) and stop for good once a link counter saturates.
__device__ __forceinline__ void payload_update (Payload *p, u64 s0, u64 s1, u32 left, u32 right, u64 ord)
{
	{
		const u64 o = min (s0 >> 24, ord);
		const u64 w0 = (o << 24) | sat_inc24 ((u32) s0 & 0xFFFFFFu, left);
		const u64 w1 = ((s1 + (1ull << 32)) & 0xFFFFFFFF00000000ull) | sat_inc24 ((u32) s1 & 0xFFFFFFu, right);
		u64 o0, o1;
		if (cas128 (p, s0, s1, w0, w1, o0, o1))
			return;
		s0 = o0;
		s1 = o1;
	}
	for (;;)
	{	// ordinal minimum + left counters
		const u64 o = min (s0 >> 24, ord);
		const u64 want = (o << 24) | sat_inc24 ((u32) s0 & 0xFFFFFFu, left);
		if (want == s0)
			break;
		const u64 old = atomicCAS (&p->ordL, s0, want);
		if (old == s0)
			break;
		s0 = old;
	}
	if (right < 4)
	{
		u32 seen = (u32) s1;
		for (;;)
		{
			if (((seen >> (6 * right)) & 63) >= LINK_SAT)
				break;
			const u32 old = atomicCAS (&p->R, seen, seen + (1u << (6 * right)));
			if (old == seen)
				break;
			seen = old;
		}
	}
	red_add_u32 (&p->count, 1u);	// fire-and-forget RED last (a RED queued ahead of a CAS halves the rate)
}

// upsert of one instance; returns 1 if this call created the node.  The table must never be
// full (the host keeps the load factor below 1 before every launch).
template <int W> struct Table;

template <> struct Table<1>
{
	static __device__ __forceinline__ int upsert (Slot1 *tab, u64 cap, const Key<1> &k, u32 left, u32 right, u64 ord)
	{
		const u64 key = k.w[0];
		u64 idx = slot_of (key_hash<1> (k), cap);
		for (;;)
		{
			Slot1 *s = tab + idx;
			u64 sk, pad, s0, s1;
			int created = 0;
			ld256 (s, sk, pad, s0, s1);
			if (sk != key)
			{
				if (sk != EMPTY64)
					goto next;
				sk = atomicCAS (&s->key, EMPTY64, key);
				if (sk == EMPTY64)
					created = 1;
				else if (sk != key)
					goto next;
				s0 = PAYLOAD0_INIT;	// seeds only; a racing writer just makes the CAS retry
				s1 = 0;
			}
			payload_update (&s->p, s0, s1, left, right, ord);
			return created;
		next:
			if (++idx == cap)
				idx = 0;
		}
	}
};

template <> struct Table<2>
{
	static __device__ __forceinline__ int upsert (Slot2 *tab, u64 cap, const Key<2> &k, u32 left, u32 right, u64 ord)
	{
		u64 idx = slot_of (key_hash<2> (k), cap);
		for (;;)
		{
			Slot2 *s = tab + idx;
			u64 k0, k1, s0, s1;
			int created = 0;
			ld256 (s, k0, k1, s0, s1);
			if (k0 != k.w[0] || k1 != k.w[1])
			{
				if (k0 != EMPTY64 || k1 != EMPTY64)
					goto next;
				if (cas128 (s, EMPTY64, EMPTY64, k.w[0], k.w[1], k0, k1))
					created = 1;
				else if (k0 != k.w[0] || k1 != k.w[1])
					goto next;
				s0 = PAYLOAD0_INIT;
				s1 = 0;
			}
			payload_update (&s->p, s0, s1, left, right, ord);
			return created;
		next:
			if (++idx == cap)
				idx = 0;
		}
	}
};

template <> struct Table<4>
{
	// 4-word keys: no 256-bit CAS exists, so a new key is claimed by locking the slot's first half with a
	// 128-bit CAS (EMPTY -> LOCKED); the second half is written, a fence orders it, and the first half is
	// published last.  A reader that finds the first half published therefore finds the whole key — if it loaded
	// all four words with ONE 256-bit load it may still have an old second half (the PTX memory model treats a
	// vector access as independent scalar accesses), so a slot whose first half matches and whose second half
	// does not is looked at once more behind a fence before it is taken for another key.
	static __device__ __forceinline__ int upsert (Slot4 *tab, u64 cap, const Key<4> &k, u32 left, u32 right, u64 ord)
	{
		u64 idx = slot_of (key_hash<4> (k), cap);
		for (;;)
		{
			Slot4 *s = tab + idx;
			u64 a0, a1, b0, b1, s0 = PAYLOAD0_INIT, s1 = 0;
			int created = 0;
			ld256 (&s->key[0], a0, a1, b0, b1);
			if (a0 == EMPTY64 && a1 == LOCKED64)
			{	// another thread is publishing this slot: look again
				__nanosleep (20);
				continue;
			}
			if (a0 == EMPTY64 && a1 == EMPTY64)
			{
				if (!cas128 (&s->key[0], EMPTY64, EMPTY64, EMPTY64, LOCKED64, a0, a1))
					continue;	// lost the race: re-examine the same slot
				st128 (&s->key[2], k.w[2], k.w[3]);
				__threadfence ();
				st128 (&s->key[0], k.w[0], k.w[1]);
				created = 1;
			}
			else
			{
				if (a0 != k.w[0] || a1 != k.w[1])
					goto next;
				if (b0 != k.w[2] || b1 != k.w[3])
				{	// published first half, other second half: another key, or a second half that had not arrived yet
					__threadfence ();
					ld128 (&s->key[2], b0, b1);
					if (b0 != k.w[2] || b1 != k.w[3])
						goto next;
				}
				ld128 (&s->p, s0, s1);
			}
			payload_update (&s->p, s0, s1, left, right, ord);
			return created;
		next:
			if (++idx == cap)
				idx = 0;
		}
	}
};

}	// namespace sdt
