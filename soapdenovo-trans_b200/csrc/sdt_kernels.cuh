// sdt_kernels.cuh — the sm_100a kernels of the pregraph hashing path.
//
//   insert_reads_kernel   chop (prlHashReads.c:164-310) fused with put_kmerset (newhash.c:411-462):
//                         persistent CTAs stage tiles of 2-bit packed reads into shared memory with
//                         16-byte vector loads, every thread takes windows of the flattened
//                         (read, offset) space and upserts them into the table.  MODE 1 writes
//                         (key, meta) records into per-owner bins instead (the send side of the
//                         multi-GPU exchange; prlHashReads.c:79-88 is the reference's partition).
//   insert_records_kernel the receive side: upsert records produced by MODE 1 on any rank.
//   init / rehash / finalize / export kernels: table maintenance and the post-pass
//                         (thread_delow prlHashReads.c:844-887, thread_mark :911-967).
#pragma once
#include "sdt_device.cuh"

namespace sdt {

static constexpr int TILE_PAD = 8;	// u32 words of padding on both sides of the staged tile
static constexpr int BLOCK = 256;
static constexpr int MAX_TILE_READS = 256;

struct ReadBatch
{
	const uint8_t *packed;	// n_reads x stride_bytes, tight-string packed
	const u32 *lens;	// or nullptr: uniform_len
	const uint8_t *nmask;	// or nullptr; n_reads x mask_stride bytes, 1 bit per base
	u64 n_reads;
	u64 first_read_ordinal;
	u32 uniform_len, stride_bytes, mask_stride;
	u32 tile_reads;		// reads per shared-memory tile (multiple of 4, <= MAX_TILE_READS)
	int K;
	u32 max_read_len, maxwin;	// maxwin = max_read_len - K + 1 (ordinal = read * maxwin + window)
};

struct Counters
{
	u64 n_nodes;		// distinct keys in the table
	u64 n_instances;	// windows processed ("kmer in reads")
	u64 overflow;		// records dropped because a bin was full (MODE 1)
	u64 n_removed, n_linear;	// finalize
	u64 export_cursor;
	u64 pad[2];
	unsigned long long freq[257];	// kmerFreq histogram (thread_mark, prlHashReads.c:946-954)
};

struct Bins
{
	u64 *records;		// n_ranks x capacity x (W + 1) u64
	u64 *counts;		// n_ranks
	u64 capacity;
	u32 n_ranks;
};

__device__ __forceinline__ u32 bswap32 (u32 x) { return __byte_perm (x, 0, 0x0123); }

__device__ __forceinline__ uint4 ldg_stream (const uint4 *p)
{
	uint4 v;
	asm volatile ("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

// one window -> canonical key and the link bases in the stored orientation
template <int W, bool NMODE>
__device__ __forceinline__ void chop_window (const u32 *rd, const u32 *mk, int len, int j, int K,
					     Key<W> &key, u32 &left, u32 &right)
{
	Key<W> f, rc;
	extract_fwd<W> (rd, j + K, K, f);
	revcomp<W> (f, K, rc);
	u32 prev = 4, next = 4;
	if (j > 0)
		prev = base_at (rd, j - 1);
	if (j + K < len)
		next = base_at (rd, j + K);
	if constexpr (NMODE)
	{	// -n: a window that holds an N becomes key 0 without links; an N neighbour is no link
		bool bad = false;
		for (int q = j >> 5; q <= (j + K - 1) >> 5; q++)
		{
			u32 m = mk[q];
			const int lo = q << 5;
			if (j > lo)
				m &= 0xFFFFFFFFu >> (j - lo);
			if (j + K < lo + 32)
				m &= 0xFFFFFFFFu << (lo + 32 - (j + K));
			bad |= (m != 0);
		}
		if (j > 0 && ((mk[(j - 1) >> 5] >> (31 - ((j - 1) & 31))) & 1))
			prev = 4;
		if (j + K < len && ((mk[(j + K) >> 5] >> (31 - ((j + K) & 31))) & 1))
			next = 4;
		if (bad)
		{
#pragma unroll
			for (int i = 0; i < W; i++)
				key.w[i] = 0;
			left = right = 4;
			return;
		}
	}
	if (key_less<W> (f, rc))
	{
		key = f;
		left = prev;
		right = next;
	}
	else
	{
		key = rc;
		left = next < 4 ? (next ^ 2u) : 4u;
		right = prev < 4 ? (prev ^ 2u) : 4u;
	}
}

template <int W> __device__ __forceinline__ void emit_record (const Bins &bins, const Key<W> &key, u32 left, u32 right, u64 ord, u64 h)
{
	const u32 owner = owner_of (h, bins.n_ranks);
	// warp-aggregated append: one atomic per (warp, destination)
	const unsigned active = __activemask ();
	const unsigned peers = __match_any_sync (active, owner);
	const int leader = __ffs (peers) - 1;
	const int lane = threadIdx.x & 31;
	u64 base = 0;
	if (lane == leader)
		base = atomicAdd (bins.counts + owner, (u64) __popc (peers));
	base = __shfl_sync (peers, base, leader);
	const u64 pos = base + __popc (peers & ((1u << lane) - 1u));
	if (pos >= bins.capacity)
		return;	// the host sees counts[owner] > capacity and reports SDTGPU_ERANGE
	u64 *rec = bins.records + ((u64) owner * bins.capacity + pos) * (W + 1);
	const u64 meta = (ord << 8) | (left << 4) | right;
	if constexpr (W == 1)
		*reinterpret_cast<ulonglong2 *> (rec) = make_ulonglong2 (key.w[0], meta);
	else
	{
#pragma unroll
		for (int i = 0; i < W; i++)
			rec[i] = key.w[i];
		rec[W] = meta;
	}
}

template <int W, bool NMODE, int MODE>
__global__ void __launch_bounds__ (BLOCK)
insert_reads_kernel (typename SlotOf<W>::type *table, u64 cap, ReadBatch rb, Bins bins, Counters *ctr)
{
	extern __shared__ __align__(16) u32 smem[];
	const int tid = threadIdx.x;
	const u32 sw = rb.stride_bytes >> 2;			// u32 words per read
	const u32 mw = NMODE ? ((rb.mask_stride + 3) >> 2) : 0;	// mask words per read
	u32 *tile = smem + TILE_PAD;
	u32 *prefix = smem + TILE_PAD + rb.tile_reads * sw + TILE_PAD;	// tile_reads + 1 entries
	u32 *mtile = prefix + rb.tile_reads + 4;
	__shared__ u32 warp_sums[BLOCK / 32];
	const u64 n_tiles = (rb.n_reads + rb.tile_reads - 1) / rb.tile_reads;
	const int K = rb.K;
	u32 created = 0;
	u64 instances = 0;

	if (tid < TILE_PAD)
	{
		smem[tid] = 0;
		smem[TILE_PAD + rb.tile_reads * sw + tid] = 0;
	}

	for (u64 t = blockIdx.x; t < n_tiles; t += gridDim.x)
	{
		const u64 r0 = t * rb.tile_reads;
		const u32 nr = (u32) min ((u64) rb.tile_reads, rb.n_reads - r0);
		// ---- stage the tile: coalesced 16-byte loads, byte-swapped so that base order == bit order
		{
			const u32 words = nr * sw;
			const uint4 *src = reinterpret_cast<const uint4 *> (rb.packed + r0 * rb.stride_bytes);
			const u32 nvec = words >> 2;
			for (u32 v = tid; v < nvec; v += BLOCK)
			{
				uint4 x = ldg_stream (src + v);
				tile[4 * v + 0] = bswap32 (x.x);
				tile[4 * v + 1] = bswap32 (x.y);
				tile[4 * v + 2] = bswap32 (x.z);
				tile[4 * v + 3] = bswap32 (x.w);
			}
			const u32 *src32 = reinterpret_cast<const u32 *> (src);
			for (u32 v = 4 * nvec + tid; v < words; v += BLOCK)
				tile[v] = bswap32 (__ldg (src32 + v));
			if constexpr (NMODE)
			{
				const uint8_t *msrc = rb.nmask + r0 * rb.mask_stride;
				for (u32 v = tid; v < nr * mw; v += BLOCK)
				{
					const u32 rr = v / mw, q = v - rr * mw;
					u32 x = 0;
					for (u32 b = 0; b < 4; b++)
					{
						const u32 byte = 4 * q + b;
						if (byte < rb.mask_stride)
							x |= (u32) msrc[(u64) rr * rb.mask_stride + byte] << (24 - 8 * b);
					}
					mtile[v] = x;
				}
			}
		}
		// ---- windows per read -> exclusive prefix (the reference's indexArray, prlHashReads.c:516-518)
		u32 total;
		const bool uniform = (rb.lens == nullptr);
		u32 nwin_u = 0;
		if (uniform)
		{
			const u32 len = min (rb.uniform_len, rb.max_read_len);
			nwin_u = len >= (u32) K + 1 ? len - K + 1 : 0;	// reads shorter than K+1 are skipped (:507)
			total = nwin_u * nr;
		}
		else
		{
			u32 c = 0;
			if ((u32) tid < nr)
			{
				const u32 len = min (rb.lens[r0 + tid], rb.max_read_len);
				c = len >= (u32) K + 1 ? len - K + 1 : 0;
			}
			u32 x = c;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1)
			{
				const u32 y = __shfl_up_sync (0xFFFFFFFFu, x, d);
				if ((tid & 31) >= d)
					x += y;
			}
			if ((tid & 31) == 31)
				warp_sums[tid >> 5] = x;
			__syncthreads ();
			u32 off = 0;
			for (int w = 0; w < (tid >> 5); w++)
				off += warp_sums[w];
			if ((u32) tid < rb.tile_reads)
				prefix[tid] = off + x - c;
			if (tid == BLOCK - 1)
				prefix[rb.tile_reads] = off + x;
			__syncthreads ();
			total = prefix[rb.tile_reads];
		}
		__syncthreads ();
		instances += (tid == 0) ? total : 0;

		for (u32 w = tid; w < total; w += BLOCK)
		{
			u32 r, j, len;
			if (uniform)
			{
				r = w / nwin_u;
				j = w - r * nwin_u;
				len = nwin_u + K - 1;
			}
			else
			{	// largest r with prefix[r] <= w
				u32 lo = 0, hi = nr - 1;
				while (lo < hi)
				{
					const u32 mid = (lo + hi + 1) >> 1;
					if (prefix[mid] <= w)
						lo = mid;
					else
						hi = mid - 1;
				}
				r = lo;
				j = w - prefix[r];
				len = prefix[r + 1] - prefix[r] + K - 1;
			}
			Key<W> key;
			u32 left, right;
			chop_window<W, NMODE> (tile + r * sw, NMODE ? (mtile + r * mw) : nullptr, (int) len, (int) j, K, key, left, right);
			const u64 ord = (rb.first_read_ordinal + r0 + r) * rb.maxwin + j;
			if (MODE == 0)
				created += Table<W>::upsert (table, cap, key, left, right, ord);
			else
				emit_record<W> (bins, key, left, right, ord, key_hash<W> (key));
		}
		__syncthreads ();	// the tile is overwritten by the next iteration
	}
	// ---- counters: one atomic per warp
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		created += __shfl_down_sync (0xFFFFFFFFu, created, d);
	if ((tid & 31) == 0 && created)
		atomicAdd (&ctr->n_nodes, (u64) created);
	if (MODE == 0 && tid == 0 && instances)	// bucketed instances are counted by the rank that inserts them
		atomicAdd (&ctr->n_instances, instances);
}

template <int W>
__global__ void __launch_bounds__ (BLOCK)
insert_records_kernel (typename SlotOf<W>::type *table, u64 cap, const u64 *records, u64 n, Counters *ctr)
{
	u32 created = 0;
	for (u64 i = blockIdx.x * (u64) BLOCK + threadIdx.x; i < n; i += (u64) gridDim.x * BLOCK)
	{
		Key<W> key;
		u64 meta;
		const u64 *rec = records + i * (W + 1);
		if constexpr (W == 1)
		{
			const ulonglong2 v = *reinterpret_cast<const ulonglong2 *> (rec);
			key.w[0] = v.x;
			meta = v.y;
		}
		else
		{
#pragma unroll
			for (int q = 0; q < W; q++)
				key.w[q] = rec[q];
			meta = rec[W];
		}
		created += Table<W>::upsert (table, cap, key, (u32) (meta >> 4) & 15u, (u32) meta & 15u, meta >> 8);
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		created += __shfl_down_sync (0xFFFFFFFFu, created, d);
	if ((threadIdx.x & 31) == 0)
	{
		if (created)
			atomicAdd (&ctr->n_nodes, (u64) created);
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
		atomicAdd (&ctr->n_instances, n);
}

// ------------------------------------------------------------------------------------------------
// slot accessors shared by init / rehash / finalize / export
template <int W> struct SlotIO;

template <> struct SlotIO<1>
{
	static __device__ __forceinline__ void init (Slot1 *s) { *reinterpret_cast<uint4 *> (s) = make_uint4 (~0u, ~0u, 0, 0); *(reinterpret_cast<uint4 *> (s) + 1) = make_uint4 (~0u, ~0u, 0, 0); }
	static __device__ __forceinline__ bool occupied (const Slot1 *s) { return s->key != EMPTY64; }
	static __device__ __forceinline__ void get (const Slot1 *s, Key<1> &k, u32 &L, u32 &R, u32 &count, u64 &ord)
	{
		k.w[0] = s->key; L = (u32) (s->links & 0xFFFFFF); R = (u32) ((s->links >> 24) & 0xFFFFFF); count = s->count; ord = s->ord;
	}
	static __device__ __forceinline__ void set_links (Slot1 *s, u32 L, u32 R) { s->links = (u64) L | ((u64) R << 24); }
	static __device__ __forceinline__ void put (Slot1 *s, const Key<1> &k, u32 L, u32 R, u32 count, u64 ord)
	{
		s->links = (u64) L | ((u64) R << 24); s->ord = ord; s->count = count;
	}
	static __device__ __forceinline__ bool claim (Slot1 *s, const Key<1> &k) { return atomicCAS (&s->key, EMPTY64, k.w[0]) == EMPTY64; }
};

template <> struct SlotIO<2>
{
	static __device__ __forceinline__ void init (Slot2 *s)
	{
		*reinterpret_cast<uint4 *> (s) = make_uint4 (~0u, ~0u, ~0u, ~0u);
		const u64 ordL = ORD40_NONE << 24;
		*(reinterpret_cast<uint4 *> (s) + 1) = make_uint4 ((u32) ordL, (u32) (ordL >> 32), 0, 0);
	}
	static __device__ __forceinline__ bool occupied (const Slot2 *s) { return !(s->key[0] == EMPTY64 && s->key[1] == EMPTY64); }
	static __device__ __forceinline__ void get (const Slot2 *s, Key<2> &k, u32 &L, u32 &R, u32 &count, u64 &ord)
	{
		k.w[0] = s->key[0]; k.w[1] = s->key[1]; L = (u32) (s->ordL & 0xFFFFFF); R = s->R & 0xFFFFFF; count = s->count; ord = s->ordL >> 24;
	}
	static __device__ __forceinline__ void set_links (Slot2 *s, u32 L, u32 R) { s->ordL = (s->ordL & ~0xFFFFFFull) | L; s->R = R; }
	static __device__ __forceinline__ void put (Slot2 *s, const Key<2> &k, u32 L, u32 R, u32 count, u64 ord)
	{
		s->ordL = (ord << 24) | L; s->R = R; s->count = count;
	}
	static __device__ __forceinline__ bool claim (Slot2 *s, const Key<2> &k)
	{
		u64 a, b;
		return cas128 (s, EMPTY64, EMPTY64, k.w[0], k.w[1], a, b);
	}
};

template <> struct SlotIO<4>
{
	static __device__ __forceinline__ void init (Slot4 *s)
	{
		uint4 *p = reinterpret_cast<uint4 *> (s);
		p[0] = make_uint4 (~0u, ~0u, ~0u, ~0u);
		p[1] = make_uint4 (~0u, ~0u, ~0u, ~0u);
		p[2] = make_uint4 (0, 0, ~0u, ~0u);
		p[3] = make_uint4 (0, 0, 0, 0);
	}
	static __device__ __forceinline__ bool occupied (const Slot4 *s) { return !(s->key[0] == EMPTY64 && s->key[1] == EMPTY64); }
	static __device__ __forceinline__ void get (const Slot4 *s, Key<4> &k, u32 &L, u32 &R, u32 &count, u64 &ord)
	{
#pragma unroll
		for (int i = 0; i < 4; i++)
			k.w[i] = s->key[i];
		L = (u32) (s->links & 0xFFFFFF); R = (u32) ((s->links >> 24) & 0xFFFFFF); count = s->count; ord = s->ord;
	}
	static __device__ __forceinline__ void set_links (Slot4 *s, u32 L, u32 R) { s->links = (u64) L | ((u64) R << 24); }
	static __device__ __forceinline__ void put (Slot4 *s, const Key<4> &k, u32 L, u32 R, u32 count, u64 ord)
	{
		s->key[2] = k.w[2]; s->key[3] = k.w[3];
		s->links = (u64) L | ((u64) R << 24); s->ord = ord; s->count = count;
	}
	static __device__ __forceinline__ bool claim (Slot4 *s, const Key<4> &k)
	{	// rehash only: keys are unique, so claiming the first half is enough (put() writes the rest)
		u64 a, b;
		return cas128 (&s->key[0], EMPTY64, EMPTY64, k.w[0], k.w[1], a, b);
	}
};

template <int W>
__global__ void __launch_bounds__ (BLOCK)
init_table_kernel (typename SlotOf<W>::type *table, u64 cap)
{
	for (u64 i = blockIdx.x * (u64) BLOCK + threadIdx.x; i < cap; i += (u64) gridDim.x * BLOCK)
		SlotIO<W>::init (table + i);
}

// move every node of `old` into the (larger, empty) table `neu`
template <int W>
__global__ void __launch_bounds__ (BLOCK)
rehash_kernel (const typename SlotOf<W>::type *old, u64 old_cap, typename SlotOf<W>::type *neu, u64 cap)
{
	for (u64 i = blockIdx.x * (u64) BLOCK + threadIdx.x; i < old_cap; i += (u64) gridDim.x * BLOCK)
	{
		if (!SlotIO<W>::occupied (old + i))
			continue;
		Key<W> k;
		u32 L, R, count;
		u64 ord;
		SlotIO<W>::get (old + i, k, L, R, count, ord);
		u64 idx = slot_of (key_hash<W> (k), cap);
		while (!SlotIO<W>::claim (neu + idx, k))
			if (++idx == cap)
				idx = 0;
		SlotIO<W>::put (neu + idx, k, L, R, count, ord);
	}
}

// thread_delow (prlHashReads.c:844-887) + thread_mark (:911-967) in one streaming pass
template <int W>
__global__ void __launch_bounds__ (BLOCK)
finalize_kernel (typename SlotOf<W>::type *table, u64 cap, int deLowKmer, Counters *ctr)
{
	__shared__ u32 hist[257];
	__shared__ u32 s_removed, s_linear;
	for (int i = threadIdx.x; i < 257; i += BLOCK)
		hist[i] = 0;
	if (threadIdx.x == 0)
		s_removed = s_linear = 0;
	__syncthreads ();
	for (u64 i = blockIdx.x * (u64) BLOCK + threadIdx.x; i < cap; i += (u64) gridDim.x * BLOCK)
	{
		if (!SlotIO<W>::occupied (table + i))
			continue;
		Key<W> k;
		u32 L, R, count;
		u64 ord;
		SlotIO<W>::get (table + i, k, L, R, count, ord);
		if (deLowKmer > 0)
		{
			u32 nL = L, nR = R;
#pragma unroll
			for (int b = 0; b < 4; b++)
			{
				const u32 cl = (L >> (6 * b)) & 63, cr = (R >> (6 * b)) & 63;
				if (cl > 0 && cl <= (u32) deLowKmer)
					nL &= ~(63u << (6 * b));
				if (cr > 0 && cr <= (u32) deLowKmer)
					nR &= ~(63u << (6 * b));
			}
			if (nL != L || nR != R)
				SlotIO<W>::set_links (table + i, nL, nR);
			L = nL;
			R = nR;
			if (L == 0 && R == 0)
				atomicAdd (&s_removed, 1u);
		}
		u32 in_num = 0, out_num = 0, l_cvg = 0, r_cvg = 0;
#pragma unroll
		for (int b = 0; b < 4; b++)
		{
			const u32 cl = (L >> (6 * b)) & 63, cr = (R >> (6 * b)) & 63;
			in_num += cl > 0; l_cvg += cl;
			out_num += cr > 0; r_cvg += cr;
		}
		atomicAdd (&hist[count == 1 ? 1 : max (l_cvg, r_cvg)], 1u);	// single == (count == 1)
		if (in_num == 1 && out_num == 1)
			atomicAdd (&s_linear, 1u);
	}
	__syncthreads ();
	for (int i = threadIdx.x; i < 257; i += BLOCK)
		if (hist[i])
			atomicAdd (&ctr->freq[i], (unsigned long long) hist[i]);
	if (threadIdx.x == 0)
	{
		if (s_removed)
			atomicAdd (&ctr->n_removed, (u64) s_removed);
		if (s_linear)
			atomicAdd (&ctr->n_linear, (u64) s_linear);
	}
}

}	// namespace sdt
