// sdt_sliced.cuh — shared pieces of the sliced build (sdt_skm.cuh): the shared-memory read tile
// (stage, flatten, locate a window) and the exclusive scan of a per-chain count.
//
// Why a sliced build at all: on B200 a request to a cold line of a larger-than-L2 table completes at
// 36.65 G/s whatever its width (profiles/r1_random_access_findings.md), so a single-pass insert (one
// load + one atomic per instance) cannot exceed 18.3 G instances/s.  Streaming traffic has no such
// limit: the sliced build turns every table access into shared-memory traffic plus sequential DRAM
// streams.  (The first version of it, which moved one 16-byte (key, meta) record per instance
// through two levels of shared-memory counting sorts, lived in this file; it reached 18.5 G/s and
// was replaced by the super-k-mer pipeline of sdt_skm.cuh.)
#pragma once
#include "sdt_kernels.cuh"

namespace sdt {

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
			if (cur2)
				cur2[base + i] = run;
		}
		run += c[i];
	}
	if (blockIdx.x == gridDim.x - 1 && tid == SCAN_NT - 1)
		off[n] = run;
}

}	// namespace sdt
