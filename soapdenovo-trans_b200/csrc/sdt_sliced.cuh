// sdt_sliced.cuh — the sliced build: put_kmerset (newhash.c:411-462) without random DRAM access.
//
// Why: on B200 a request to a cold line of a larger-than-L2 table completes at 36.65 G/s whatever
// its width (profiles/r1_random_access_findings.md), so a single-pass insert (one load + one atomic
// per instance) cannot exceed 18.3 G instances/s.  Streaming traffic has no such limit.  The sliced
// build therefore turns every table access into shared-memory traffic plus sequential DRAM streams:
//
//   the table is n_slices x slice_slots slots; a key lives in slice  mulhi64(mix(key), n_slices),
//   probing stays inside the slice (home = mulhi32(low32(mix), slice_slots), +1 with wrap);
//
//   slice_count_kernel     chop (prlHashReads.c:164-310) every window, histogram of instances per
//                          slice (RED into an L2-resident array) — sizes every stream exactly;
//   slice_scan_*           exclusive scan -> record offset of every slice and level-1 partition;
//   slice_scatter1_kernel  chop again, (key, meta) records into P1 level-1 partitions (P2 slices
//                          each); a tile of records is ordered by partition in shared memory first,
//                          so the global stores are runs of consecutive records;
//   slice_scatter2_kernel  one level-1 partition at a time: records into their slice's run;
//   slice_build_kernel     one CTA per slice: the slice's table image is built in shared memory
//                          (key claim by 64-bit CAS or a per-slot lock word, 32-bit counters and the
//                          ordinal minimum with shared-memory atomics, link counters clamped to the
//                          reference's 6-bit saturation when written out) and streamed to the table
//                          in the ordinary slot layout, empty slots included — so finalize / export /
//                          checksum run unchanged and no table initialisation pass is needed.
//
// All updates commute, so the result is bit-identical to the reference's sequential put_kmerset.
#pragma once
#include "sdt_kernels.cuh"

namespace sdt {

struct SliceGeom
{
	u32 n_slices;		// table capacity = n_slices * slice_slots
	u32 slice_slots;	// S
	u32 P1, P2;		// level-1 partitions, slices per level-1 partition (P1 = ceil (n_slices / P2))
};

__device__ __forceinline__ u32 slice_of (u64 h, u32 n_slices) { return (u32) __umul64hi (h, (u64) n_slices); }
__device__ __forceinline__ u32 home_of (u64 h, u32 S) { return __umulhi ((u32) h, S); }

static constexpr int CNT_NT = 256;	// slice_count_kernel
static constexpr int SC_NT = 512;	// scatter kernels
static constexpr int BD_NT = 1024;	// slice_build_kernel
template <int W> struct ScatterCfg { static constexpr int RPT = W == 4 ? 4 : 8; };	// records per thread and tile
static constexpr u32 NO_BIN = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------------
// a tile of 2-bit packed reads in shared memory (same layout as insert_reads_kernel's)
template <bool NMODE> struct ReadTile
{
	u32 *tile, *prefix, *mtile;
	u32 sw, mw;		// u32 words per read: packed bases, N mask
	u32 nr, total, nwin_u;	// reads in the tile, windows in the tile, windows per read (uniform lengths)
	bool uniform;
	u64 r0;			// first read of the tile
};

__host__ __device__ inline size_t tile_words (const ReadBatch &rb, bool nmode)
{
	const size_t sw = rb.stride_bytes / 4, mw = nmode ? (rb.mask_stride + 3) / 4 : 0;
	const size_t w = 2 * TILE_PAD + rb.tile_reads * sw + rb.tile_reads + 4 + rb.tile_reads * mw;
	return (w + 3) & ~(size_t) 3;	// what follows stays 16-byte aligned
}

template <bool NMODE>
__device__ __forceinline__ void tile_setup (ReadTile<NMODE> &rt, u32 *smem, const ReadBatch &rb)
{
	rt.sw = rb.stride_bytes >> 2;
	rt.mw = NMODE ? ((rb.mask_stride + 3) >> 2) : 0;
	rt.tile = smem + TILE_PAD;
	rt.prefix = smem + TILE_PAD + rb.tile_reads * rt.sw + TILE_PAD;
	rt.mtile = rt.prefix + rb.tile_reads + 4;
	if (threadIdx.x < TILE_PAD)
	{
		smem[threadIdx.x] = 0;
		smem[TILE_PAD + rb.tile_reads * rt.sw + threadIdx.x] = 0;
	}
}

// stage tile t (coalesced 16-byte loads, byte-swapped so that base order == bit order) and flatten
// its windows (the reference's indexArray, prlHashReads.c:516-518).  Ends with a barrier.
template <bool NMODE, int NT>
__device__ __forceinline__ void tile_stage (ReadTile<NMODE> &rt, const ReadBatch &rb, u64 t, u32 *warp_sums)
{
	const u32 tid = threadIdx.x;
	const u32 K = (u32) rb.K;
	rt.r0 = t * rb.tile_reads;
	rt.nr = (u32) min ((u64) rb.tile_reads, rb.n_reads - rt.r0);
	{
		const u32 words = rt.nr * rt.sw;
		const uint4 *src = reinterpret_cast<const uint4 *> (rb.packed + rt.r0 * rb.stride_bytes);
		const u32 nvec = words >> 2;
		for (u32 v = tid; v < nvec; v += NT)
		{
			const uint4 x = ldg_stream (src + v);
			rt.tile[4 * v + 0] = bswap32 (x.x);
			rt.tile[4 * v + 1] = bswap32 (x.y);
			rt.tile[4 * v + 2] = bswap32 (x.z);
			rt.tile[4 * v + 3] = bswap32 (x.w);
		}
		const u32 *src32 = reinterpret_cast<const u32 *> (src);
		for (u32 v = 4 * nvec + tid; v < words; v += NT)
			rt.tile[v] = bswap32 (__ldg (src32 + v));
		if constexpr (NMODE)
		{
			const uint8_t *msrc = rb.nmask + rt.r0 * rb.mask_stride;
			for (u32 v = tid; v < rt.nr * rt.mw; v += NT)
			{
				const u32 rr = v / rt.mw, q = v - rr * rt.mw;
				u32 x = 0;
				for (u32 b = 0; b < 4; b++)
				{
					const u32 byte = 4 * q + b;
					if (byte < rb.mask_stride)
						x |= (u32) msrc[(u64) rr * rb.mask_stride + byte] << (24 - 8 * b);
				}
				rt.mtile[v] = x;
			}
		}
	}
	rt.uniform = (rb.lens == nullptr);
	rt.nwin_u = 0;
	if (rt.uniform)
	{
		const u32 len = min (rb.uniform_len, rb.max_read_len);
		rt.nwin_u = len >= K + 1 ? len - K + 1 : 0;	// reads shorter than K+1 are skipped (prlHashReads.c:507)
		rt.total = rt.nwin_u * rt.nr;
	}
	else
	{	// tile_reads <= MAX_TILE_READS <= NT: one read per thread
		u32 c = 0;
		if (tid < rt.nr)
		{
			const u32 len = min (rb.lens[rt.r0 + tid], rb.max_read_len);
			c = len >= K + 1 ? len - K + 1 : 0;
		}
		u32 x = c;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const u32 y = __shfl_up_sync (0xFFFFFFFFu, x, d);
			if ((tid & 31) >= (u32) d)
				x += y;
		}
		if ((tid & 31) == 31)
			warp_sums[tid >> 5] = x;
		__syncthreads ();
		u32 off = 0;
		for (u32 w = 0; w < (tid >> 5); w++)
			off += warp_sums[w];
		if (tid < rb.tile_reads)
			rt.prefix[tid] = off + x - c;
		if (tid == NT - 1)
			rt.prefix[rb.tile_reads] = off + x;
		__syncthreads ();
		rt.total = rt.prefix[rb.tile_reads];
	}
	__syncthreads ();
}

// window w of the tile -> canonical key, link bases, instance ordinal
template <int W, bool NMODE>
__device__ __forceinline__ void tile_chop (const ReadTile<NMODE> &rt, const ReadBatch &rb, u32 w,
					   Key<W> &key, u32 &left, u32 &right, u64 &ord)
{
	u32 r, j, len;
	if (rt.uniform)
	{
		r = w / rt.nwin_u;
		j = w - r * rt.nwin_u;
		len = rt.nwin_u + rb.K - 1;
	}
	else
	{	// largest r with prefix[r] <= w
		u32 lo = 0, hi = rt.nr - 1;
		while (lo < hi)
		{
			const u32 mid = (lo + hi + 1) >> 1;
			if (rt.prefix[mid] <= w)
				lo = mid;
			else
				hi = mid - 1;
		}
		r = lo;
		j = w - rt.prefix[r];
		len = rt.prefix[r + 1] - rt.prefix[r] + rb.K - 1;
	}
	chop_window<W, NMODE> (rt.tile + r * rt.sw, NMODE ? (rt.mtile + r * rt.mw) : nullptr, (int) len, (int) j, rb.K, key, left, right);
	ord = (rb.first_read_ordinal + rt.r0 + r) * rb.maxwin + j;
}

// ------------------------------------------------------------------------------------------------
// instances per slice
template <int W, bool NMODE>
__global__ void __launch_bounds__ (CNT_NT)
slice_count_kernel (ReadBatch rb, SliceGeom g, u32 *hist)
{
	extern __shared__ __align__(16) u32 smem[];
	__shared__ u32 warp_sums[CNT_NT / 32];
	ReadTile<NMODE> rt;
	tile_setup<NMODE> (rt, smem, rb);
	const u64 n_tiles = (rb.n_reads + rb.tile_reads - 1) / rb.tile_reads;
	for (u64 t = blockIdx.x; t < n_tiles; t += gridDim.x)
	{
		tile_stage<NMODE, CNT_NT> (rt, rb, t, warp_sums);
		for (u32 w = threadIdx.x; w < rt.total; w += CNT_NT)
		{
			Key<W> key;
			u32 left, right;
			u64 ord;
			tile_chop<W, NMODE> (rt, rb, w, key, left, right, ord);
			const u64 h = key_hash<W> (key);
			if (rb.owner_ranks <= 1 || owner_of (h, rb.owner_ranks) == rb.owner_rank)
				atomicAdd (hist + slice_of (h, g.n_slices), 1u);	// RED, result unused
		}
		__syncthreads ();	// the tile is overwritten by the next iteration
	}
}

// ---- exclusive scan of the histogram: 64-bit record offsets of every slice.
static constexpr int SCAN_NT = 1024, SCAN_PER = 8, SCAN_SEG = SCAN_NT * SCAN_PER;

__global__ void __launch_bounds__ (SCAN_NT)
slice_scan_sums_kernel (const u32 *hist, u32 n, u64 *seg_sum)
{
	__shared__ u64 part[SCAN_NT / 32];
	const u32 base = blockIdx.x * SCAN_SEG + threadIdx.x * SCAN_PER;
	u64 s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_PER; i++)
		if (base + i < n)
			s += hist[base + i];
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		s += __shfl_down_sync (0xFFFFFFFFu, s, d);
	if ((threadIdx.x & 31) == 0)
		part[threadIdx.x >> 5] = s;
	__syncthreads ();
	if (threadIdx.x == 0)
	{
		u64 tsum = 0;
		for (int i = 0; i < SCAN_NT / 32; i++)
			tsum += part[i];
		seg_sum[blockIdx.x] = tsum;
	}
}

// off[p] = records before slice p (off[n] = total), cur2[p] = off[p] (level-2 cursors)
__global__ void __launch_bounds__ (SCAN_NT)
slice_scan_kernel (const u32 *hist, u32 n, const u64 *seg_sum, u64 *off, u64 *cur2)
{
	__shared__ u64 part[SCAN_NT / 32];
	__shared__ u64 s_base;
	const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	{	// records before this segment
		u64 s = 0;
		for (u32 i = tid; i < blockIdx.x; i += SCAN_NT)
			s += seg_sum[i];
#pragma unroll
		for (int d = 16; d > 0; d >>= 1)
			s += __shfl_down_sync (0xFFFFFFFFu, s, d);
		if (lane == 0)
			part[wid] = s;
		__syncthreads ();
		if (tid == 0)
		{
			u64 tsum = 0;
			for (int i = 0; i < SCAN_NT / 32; i++)
				tsum += part[i];
			s_base = tsum;
		}
		__syncthreads ();
	}
	const u32 base = blockIdx.x * SCAN_SEG + tid * SCAN_PER;
	u32 c[SCAN_PER];
	u64 s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_PER; i++)
	{
		c[i] = base + i < n ? hist[base + i] : 0;
		s += c[i];
	}
	u64 x = s;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const u64 y = __shfl_up_sync (0xFFFFFFFFu, x, d);
		if (lane >= (u32) d)
			x += y;
	}
	__syncthreads ();	// part[] was read above
	if (lane == 31)
		part[wid] = x;
	__syncthreads ();
	u64 run = s_base + x - s;
	for (u32 w = 0; w < wid; w++)
		run += part[w];
#pragma unroll
	for (int i = 0; i < SCAN_PER; i++)
	{
		if (base + i < n)
		{
			off[base + i] = run;
			cur2[base + i] = run;
		}
		run += c[i];
	}
	if (blockIdx.x == gridDim.x - 1 && tid == SCAN_NT - 1)
		off[n] = run;
}

// level-1 view: off1[q] = off[min (q * P2, n_slices)] (q = 0..P1), cur1[q] = off1[q], and the
// exclusive prefix of scatter2 tiles per level-1 partition, tpre[q] (q = 0..P1).  One CTA.
__global__ void __launch_bounds__ (1024)
slice_level1_kernel (const u64 *off, SliceGeom g, u32 tile_recs, u64 *off1, u64 *cur1, u64 *tpre)
{
	__shared__ u64 part[1024];
	__shared__ u64 carry;
	if (threadIdx.x == 0)
		carry = 0;
	__syncthreads ();
	for (u32 q0 = 0; q0 < g.P1; q0 += 1024)
	{
		const u32 q = q0 + threadIdx.x;
		u64 c = 0;
		if (q < g.P1)
		{
			const u64 a = off[min ((u64) q * g.P2, (u64) g.n_slices)], b = off[min ((u64) (q + 1) * g.P2, (u64) g.n_slices)];
			off1[q] = a;
			cur1[q] = a;
			if (q == g.P1 - 1)
				off1[g.P1] = b;
			c = (b - a + tile_recs - 1) / tile_recs;
		}
		part[threadIdx.x] = c;
		__syncthreads ();
		for (int d = 1; d < 1024; d <<= 1)
		{
			const u64 y = threadIdx.x >= (u32) d ? part[threadIdx.x - d] : 0;
			__syncthreads ();
			part[threadIdx.x] += y;
			__syncthreads ();
		}
		if (q < g.P1)
			tpre[q] = carry + part[threadIdx.x] - c;
		__syncthreads ();
		if (threadIdx.x == 1023)
			carry += part[1023];
		__syncthreads ();
	}
	if (threadIdx.x == 0)
		tpre[g.P1] = carry;
}

// ------------------------------------------------------------------------------------------------
// exclusive prefix sum of cnt[0..n) (shared memory) into off[0..n); every thread gets the total.
// Contains two barriers; all threads of the block must call it.
template <int NT>
__device__ __forceinline__ u32 block_excl_scan (const u32 *cnt, u32 *off, u32 n, u32 *warp_sums)
{
	const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const u32 per = (n + NT - 1) / NT;
	const u32 b0 = min (n, tid * per), b1 = min (n, b0 + per);
	u32 s = 0;
	for (u32 b = b0; b < b1; b++)
		s += cnt[b];
	u32 x = s;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const u32 y = __shfl_up_sync (0xFFFFFFFFu, x, d);
		if (lane >= (u32) d)
			x += y;
	}
	if (lane == 31)
		warp_sums[wid] = x;
	__syncthreads ();
	u32 run = x - s, total = 0;
	for (u32 w = 0; w < NT / 32; w++)
	{
		const u32 v = warp_sums[w];
		if (w < wid)
			run += v;
		total += v;
	}
	for (u32 b = b0; b < b1; b++)
	{
		const u32 c = cnt[b];
		off[b] = run;
		run += c;
	}
	__syncthreads ();
	return total;
}

struct ScatterSmem
{
	u64 *staging;		// tile_recs x (W + 1) u64
	u64 *gbase;		// [bins] destination record index of the tile's run in each bin
	u32 *cnt, *soff;	// [bins]
	unsigned short *bin16;	// [tile_recs]
	u32 *warp_sums;		// [NT / 32]
};

__host__ __device__ inline size_t scatter_smem_bytes (int W, u32 tile_recs, u32 max_bins)
{
	return (size_t) tile_recs * 8 * (W + 1) + (size_t) max_bins * 16 + (size_t) tile_recs * 2 + 16;
}

__device__ __forceinline__ void scatter_smem_setup (ScatterSmem &ss, void *base, int W, u32 tile_recs, u32 max_bins, u32 *warp_sums)
{
	ss.staging = static_cast<u64 *> (base);
	ss.gbase = ss.staging + (size_t) tile_recs * (W + 1);
	ss.cnt = reinterpret_cast<u32 *> (ss.gbase + max_bins);
	ss.soff = ss.cnt + max_bins;
	ss.bin16 = reinterpret_cast<unsigned short *> (ss.soff + max_bins);
	ss.warp_sums = warp_sums;
}

// One tile of records held in registers (RPT per thread; bin NO_BIN = no record) -> `out`, grouped
// by bin: count per bin in shared memory (the atomic's return value is the record's rank inside
// its bin), scan, ONE global atomic per non-empty bin reserves the tile's run in that bin, the
// records are ordered by bin in shared memory and copied out so that consecutive threads write
// consecutive records.  cursors[b] holds absolute record indices; out_base is subtracted (the
// level-2 output is a scratch area that starts at the group's first record).
template <int W, int RPT, int NT>
__device__ __forceinline__ void scatter_tile (const Key<W> (&key)[RPT], const u64 (&meta)[RPT], const u32 (&bin)[RPT],
					      u32 bins, u64 *cursors, u64 out_base, u64 *out, const ScatterSmem &ss)
{
	constexpr int REC = W + 1;
	const u32 tid = threadIdx.x;
	for (u32 b = tid; b < bins; b += NT)
		ss.cnt[b] = 0;
	__syncthreads ();
	u32 rank[RPT];
#pragma unroll
	for (int k = 0; k < RPT; k++)
		rank[k] = bin[k] != NO_BIN ? atomicAdd (&ss.cnt[bin[k]], 1u) : 0u;
	__syncthreads ();
	const u32 n = block_excl_scan<NT> (ss.cnt, ss.soff, bins, ss.warp_sums);
	for (u32 b = tid; b < bins; b += NT)
	{
		const u32 c = ss.cnt[b];
		if (c)
			ss.gbase[b] = atomicAdd (cursors + b, (u64) c) - out_base;
	}
#pragma unroll
	for (int k = 0; k < RPT; k++)
		if (bin[k] != NO_BIN)
		{
			const u32 pos = ss.soff[bin[k]] + rank[k];
			if constexpr (W == 1)
				reinterpret_cast<ulonglong2 *> (ss.staging)[pos] = make_ulonglong2 (key[k].w[0], meta[k]);
			else
			{
#pragma unroll
				for (int q = 0; q < W; q++)
					ss.staging[(size_t) pos * REC + q] = key[k].w[q];
				ss.staging[(size_t) pos * REC + W] = meta[k];
			}
			ss.bin16[pos] = (unsigned short) bin[k];
		}
	__syncthreads ();
	if constexpr (W == 1)
	{
		ulonglong2 *o = reinterpret_cast<ulonglong2 *> (out);
		const ulonglong2 *s = reinterpret_cast<const ulonglong2 *> (ss.staging);
		for (u32 i = tid; i < n; i += NT)
		{
			const u32 b = ss.bin16[i];
			o[ss.gbase[b] + (i - ss.soff[b])] = s[i];
		}
	}
	else
	{
		for (u32 q = tid; q < n * REC; q += NT)
		{
			const u32 i = q / REC, wd = q - i * REC;
			const u32 b = ss.bin16[i];
			out[(ss.gbase[b] + (i - ss.soff[b])) * REC + wd] = ss.staging[q];
		}
	}
	__syncthreads ();	// staging and counters are reused by the next tile
}

// level 1: reads -> records in P1 partitions (cur1[q] starts at off1[q])
template <int W, bool NMODE>
__global__ void __launch_bounds__ (SC_NT, 2)
slice_scatter1_kernel (ReadBatch rb, SliceGeom g, u64 *cur1, u64 *rec1)
{
	constexpr int RPT = ScatterCfg<W>::RPT;
	constexpr u32 TILE_RECS = SC_NT * RPT;
	extern __shared__ __align__(16) u32 smem[];
	__shared__ u32 warp_sums[SC_NT / 32];
	ReadTile<NMODE> rt;
	tile_setup<NMODE> (rt, smem, rb);
	ScatterSmem ss;
	scatter_smem_setup (ss, smem + tile_words (rb, NMODE), W, TILE_RECS, g.P1, warp_sums);
	const u64 n_tiles = (rb.n_reads + rb.tile_reads - 1) / rb.tile_reads;
	for (u64 t = blockIdx.x; t < n_tiles; t += gridDim.x)
	{
		tile_stage<NMODE, SC_NT> (rt, rb, t, warp_sums);
		for (u32 c0 = 0; c0 < rt.total; c0 += TILE_RECS)
		{
			Key<W> key[RPT];
			u64 meta[RPT];
			u32 bin[RPT];
#pragma unroll
			for (int k = 0; k < RPT; k++)
			{
				const u32 w = c0 + k * SC_NT + threadIdx.x;
				bin[k] = NO_BIN;
				meta[k] = 0;
#pragma unroll
				for (int q = 0; q < W; q++)
					key[k].w[q] = 0;
				if (w < rt.total)
				{
					u32 left, right;
					u64 ord;
					tile_chop<W, NMODE> (rt, rb, w, key[k], left, right, ord);
					const u64 h = key_hash<W> (key[k]);
					if (rb.owner_ranks <= 1 || owner_of (h, rb.owner_ranks) == rb.owner_rank)
					{
						bin[k] = slice_of (h, g.n_slices) / g.P2;
						meta[k] = (ord << 8) | (left << 4) | right;
					}
				}
			}
			scatter_tile<W, RPT, SC_NT> (key, meta, bin, g.P1, cur1, 0, rec1, ss);
		}
		__syncthreads ();	// the read tile is overwritten by the next iteration
	}
}

// level 2: the records of level-1 partitions [q_lo, q_hi) -> the run of their slice, in the scratch
// area rec2 (which starts at record out_base = off1[q_lo]).  Tiles never straddle a level-1
// partition, so a tile has at most P2 bins.
template <int W>
__global__ void __launch_bounds__ (SC_NT, 2)
slice_scatter2_kernel (const u64 *rec1, const u64 *off1, const u64 *tpre, u32 q_lo, u32 q_hi, SliceGeom g,
		       u64 *cur2, u64 *rec2, u64 out_base)
{
	constexpr int RPT = ScatterCfg<W>::RPT;
	constexpr u32 TILE_RECS = SC_NT * RPT;
	constexpr int REC = W + 1;
	extern __shared__ __align__(16) u32 smem[];
	__shared__ u32 warp_sums[SC_NT / 32];
	ScatterSmem ss;
	scatter_smem_setup (ss, smem, W, TILE_RECS, g.P2, warp_sums);
	const u64 t_begin = tpre[q_lo], t_end = tpre[q_hi];
	for (u64 t = t_begin + blockIdx.x; t < t_end; t += gridDim.x)
	{
		u32 lo = q_lo, hi = q_hi - 1;	// largest q with tpre[q] <= t
		while (lo < hi)
		{
			const u32 mid = (lo + hi + 1) >> 1;
			if (tpre[mid] <= t)
				lo = mid;
			else
				hi = mid - 1;
		}
		const u32 q = lo;
		const u64 first = off1[q] + (t - tpre[q]) * TILE_RECS, last = min (first + TILE_RECS, off1[q + 1]);
		const u32 pbase = q * g.P2, bins = min (g.P2, g.n_slices - pbase);
		Key<W> key[RPT];
		u64 meta[RPT];
		u32 bin[RPT];
#pragma unroll
		for (int k = 0; k < RPT; k++)
		{
			const u64 i = first + k * SC_NT + threadIdx.x;
			bin[k] = NO_BIN;
			meta[k] = 0;
#pragma unroll
			for (int w = 0; w < W; w++)
				key[k].w[w] = 0;
			if (i < last)
			{
				load_record<W> (rec1 + i * REC, key[k], meta[k]);
				bin[k] = slice_of (key_hash<W> (key[k]), g.n_slices) - pbase;
			}
		}
		scatter_tile<W, RPT, SC_NT> (key, meta, bin, bins, cur2 + pbase, out_base, rec2, ss);
	}
}

// ------------------------------------------------------------------------------------------------
// The slice's table image in shared memory.  Shared-memory atomics are the scarce resource here
// (2 cycles per lane on this part, 20x the cost of a load), so an instance costs exactly ONE of
// them: every slot has a 5 x 5 matrix of plain 32-bit cells indexed by (left, right) with 4 = "no
// neighbour" — the reference's update_kmer (newhash.c:71-96) touches count, one left and one right
// counter per instance, and all three are sums over that matrix:
//     count = sum of all cells (mod 2^32),  L[b] = min (63, sum of row b),  R[b] = min (63, sum of column b).
// The ordinal minimum needs an atomic only when it improves (a load otherwise).
static constexpr int CELLS = 25;

template <int W> struct SliceImage
{
	u64 *key;	// [S * W]
	u64 *ord;	// [S]
	u32 *cell;	// [CELLS * S], cell c of slot i at c * S + i
	u32 *state;	// [S] (W > 1): 0 empty, 1 key being written, 2 occupied
};

__host__ __device__ inline size_t slice_image_bytes (int W, u32 S)
{
	return (size_t) S * (8 * W + 8 + 4 * CELLS + (W > 1 ? 4 : 0));
}

// find or claim the slot of `key` inside the slice; returns the slot, or S if the slice is full
template <int W>
__device__ __forceinline__ u32 image_find (const SliceImage<W> &im, u32 S, const Key<W> &key, u32 idx, u32 &created)
{
	for (u32 tries = 0; tries < S;)
	{
		if constexpr (W == 1)
		{
			u64 k = *reinterpret_cast<volatile u64 *> (im.key + idx);
			if (k == key.w[0])
				return idx;
			if (k == EMPTY64)
			{
				k = atomicCAS (im.key + idx, EMPTY64, key.w[0]);
				if (k == EMPTY64)
				{
					created++;
					return idx;
				}
				if (k == key.w[0])
					return idx;
			}
		}
		else
		{
			const u32 st = *reinterpret_cast<volatile u32 *> (im.state + idx);
			if (st == 0u)
			{
				if (atomicCAS (im.state + idx, 0u, 1u) == 0u)
				{	// claimed: publish the key, then open the slot (no waiting inside this branch)
#pragma unroll
					for (int q = 0; q < W; q++)
						*reinterpret_cast<volatile u64 *> (im.key + (size_t) idx * W + q) = key.w[q];
					__threadfence_block ();
					*reinterpret_cast<volatile u32 *> (im.state + idx) = 2u;
					created++;
					return idx;
				}
				continue;	// lost the race: look at the same slot again
			}
			if (st == 1u)
				continue;	// its key is being written
			bool eq = true;
#pragma unroll
			for (int q = 0; q < W; q++)
				eq &= (*reinterpret_cast<volatile u64 *> (im.key + (size_t) idx * W + q) == key.w[q]);
			if (eq)
				return idx;
		}
		if (++idx == S)
			idx = 0;
		tries++;
	}
	return S;
}

// One CTA per slice: build (or, with merge != 0, update) the slice's table image in shared memory
// from its run of records and stream it to the table.  The image is empty when a slice starts (the
// write-out of the previous slice cleans up behind itself).
template <int W>
__global__ void __launch_bounds__ (BD_NT)
slice_build_kernel (typename SlotOf<W>::type *table, SliceGeom g, const u64 *rec2, const u64 *off, u64 out_base,
		    u32 p_lo, u32 p_hi, int merge, Counters *ctr)
{
	typedef typename SlotOf<W>::type S_t;
	constexpr int REC = W + 1;
	extern __shared__ __align__(16) u32 smem[];
	__shared__ u32 s_created, s_full;
	const u32 S = g.slice_slots, tid = threadIdx.x;
	SliceImage<W> im;
	im.key = reinterpret_cast<u64 *> (smem);
	im.ord = im.key + (size_t) S * W;
	im.cell = reinterpret_cast<u32 *> (im.ord + S);
	im.state = im.cell + (size_t) S * CELLS;
	u64 instances = 0, nodes = 0;	// thread 0 only
	if (tid == 0)
		s_created = s_full = 0;
	for (u32 i = tid; i < S; i += BD_NT)
	{
#pragma unroll
		for (int q = 0; q < W; q++)
			im.key[(size_t) i * W + q] = EMPTY64;
		im.ord[i] = ORD40_NONE;
		if constexpr (W > 1)
			im.state[i] = 0u;
	}
	for (u32 i = tid; i < S * CELLS; i += BD_NT)
		im.cell[i] = 0u;
	__syncthreads ();
	for (u32 p = p_lo + blockIdx.x; p < p_hi; p += gridDim.x)
	{
		S_t *slice = table + (u64) p * S;
		if (merge)
		{	// the keys this slice already holds, at their slots (their payload is added at write-out)
			for (u32 i = tid; i < S; i += BD_NT)
				if (SlotIO<W>::occupied (slice + i))
				{
#pragma unroll
					for (int q = 0; q < W; q++)
						im.key[(size_t) i * W + q] = SlotIO<W>::keyp (slice + i)[q];
					if constexpr (W > 1)
						im.state[i] = 2u;
				}
			__syncthreads ();
		}
		const u64 r0 = off[p] - out_base, r1 = off[p + 1] - out_base;
		u32 created = 0;
		bool full = false;
		for (u64 i0 = r0; i0 < r1; i0 += 2 * BD_NT)
		{	// two records per thread in flight
			Key<W> key[2];
			u64 meta[2];
			bool have[2];
#pragma unroll
			for (int u = 0; u < 2; u++)
			{
				const u64 i = i0 + u * BD_NT + tid;
				have[u] = i < r1;
				if (have[u])
					load_record<W> (rec2 + i * REC, key[u], meta[u]);
			}
#pragma unroll
			for (int u = 0; u < 2; u++)
			{
				if (!have[u])
					continue;
				const u64 h = key_hash<W> (key[u]);
				const u32 idx = image_find<W> (im, S, key[u], home_of (h, S), created);
				if (idx == S)
				{
					full = true;
					continue;
				}
				const u32 left = (u32) (meta[u] >> 4) & 15u, right = (u32) meta[u] & 15u;	// 0..4
				const u64 ord = meta[u] >> 8;
				atomicAdd (im.cell + (left * 5 + right) * S + idx, 1u);
				if (ord < *reinterpret_cast<volatile u64 *> (im.ord + idx))
					atomicMin (im.ord + idx, ord);
			}
		}
		if (created)
			atomicAdd (&s_created, created);
		if (full)
			s_full = 1;
		__syncthreads ();
		// ---- stream the image out in the table's slot layout (empty slots too) and clean it
		for (u32 i = tid; i < S; i += BD_NT)
		{
			bool occ;
			if constexpr (W == 1)
				occ = im.key[i] != EMPTY64;
			else
				occ = im.state[i] == 2u;
			u64 w0 = PAYLOAD0_INIT, w1 = 0;
			Key<W> k;
#pragma unroll
			for (int q = 0; q < W; q++)
				k.w[q] = EMPTY64;
			if (occ)
			{
				u32 row[4] = { 0, 0, 0, 0 }, col[4] = { 0, 0, 0, 0 }, count = 0;
#pragma unroll
				for (int l = 0; l < 5; l++)
#pragma unroll
					for (int r = 0; r < 5; r++)
					{
						const u32 c = im.cell[(l * 5 + r) * S + i];
						im.cell[(l * 5 + r) * S + i] = 0u;
						count += c;
						if (l < 4)
							row[l] += min (c, LINK_SAT);	// clamped terms: no 32-bit wrap, same min (63, sum)
						if (r < 4)
							col[r] += min (c, LINK_SAT);
					}
				u64 ord = im.ord[i];
				u32 oL = 0, oR = 0;
				if (merge && SlotIO<W>::occupied (slice + i))
				{	// same key (slots never move): add what the table already holds
					Key<W> ok;
					u32 ocount;
					u64 oord;
					SlotIO<W>::get (slice + i, ok, oL, oR, ocount, oord);
					count += ocount;
					ord = min (ord, oord);
				}
				u32 L = 0, R = 0;
#pragma unroll
				for (int b = 0; b < 4; b++)
				{
					L |= min (row[b] + ((oL >> (6 * b)) & 63u), LINK_SAT) << (6 * b);
					R |= min (col[b] + ((oR >> (6 * b)) & 63u), LINK_SAT) << (6 * b);
				}
				w0 = (ord << 24) | L;
				w1 = ((u64) count << 32) | R;
#pragma unroll
				for (int q = 0; q < W; q++)
				{
					k.w[q] = im.key[(size_t) i * W + q];
					im.key[(size_t) i * W + q] = EMPTY64;
				}
				im.ord[i] = ORD40_NONE;
				if constexpr (W > 1)
					im.state[i] = 0u;
			}
			if constexpr (W == 1)
				st256 (slice + i, k.w[0], occ ? 0ull : EMPTY64, w0, w1);
			else if constexpr (W == 2)
				st256 (slice + i, k.w[0], k.w[1], w0, w1);
			else
			{
				st256 (slice + i, k.w[0], k.w[1], k.w[2], k.w[3]);
				st256 (reinterpret_cast<u64 *> (slice + i) + 4, w0, w1, 0ull, 0ull);
			}
		}
		if (tid == 0)
		{	// between the two barriers nobody adds to s_created
			instances += r1 - r0;
			nodes += s_created;
			s_created = 0;
		}
		__syncthreads ();
	}
	if (tid == 0)
	{
		if (instances)
			atomicAdd (&ctr->n_instances, instances);
		if (nodes)
			atomicAdd (&ctr->n_nodes, nodes);
		if (s_full)
			atomicAdd (&ctr->overflow, 1ull);
	}
}

}	// namespace sdt
