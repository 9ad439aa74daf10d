// sdt_skm.cuh — the sliced build over super-k-mers: put_kmerset (newhash.c:411-462) without random
// DRAM access and with one tenth of the k-mer records.
//
// Why.  A random-access insert is capped by the rate at which B200 completes requests to cold lines
// (36.65 G/s, one load + one atomic per instance: 18.3 G instances/s, profiles/r1_random_access_findings.md),
// and a pipeline that moves one 16-byte record per instance through shared-memory counting sorts is
// capped by shared-memory atomics (2 cycles per lane: count + two scatter levels + build measured
// 18.5 G/s).  So instances must not travel one by one.  Consecutive windows of a read share their
// minimizer (the smallest hashed canonical m-mer inside the window) about eight at a time; a run of
// windows with the same minimizer — a super-k-mer — travels as ONE record that holds the run's
// bases, and the table is cut into slices by minimizer, so every instance of a k-mer, from either
// strand, meets in the same slice:
//
//   skm_emit_kernel     reads -> super-k-mer records (unordered stream) + records per slice (REDG);
//                       m-mer hashes and window minima live in shared memory, one pass over the reads;
//   slice_scan_*        exclusive scan of the per-slice record counts (sdt_sliced.cuh);
//   skm_scatter_kernel  every record to its slice's run (one L2-resident cursor atomic per record);
//   skm_build_kernel    one CTA per slice: the records are chopped (chop_window, the same code the
//                       single-pass path uses: prlHashReads.c:164-310) and upserted into the slice's
//                       table image in shared memory — one shared-memory atomic per instance — and the
//                       image is compacted into the node store in the ordinary slot layout, so
//                       finalize / export / checksum run unchanged over [0, n_nodes).
//
// A slice whose distinct k-mers do not fit its image writes nothing and is retried split by k-mer
// hash (work items (slice, r, R): keys with hash % R == r), so no input can break it.
// All updates commute: the result is bit-identical to the reference's sequential put_kmerset.
#pragma once
#include "sdt_sliced.cuh"
#include "sdt_chain.cuh"

namespace sdt {

struct SkmGeom
{
	u32 n_slices;		// slices of the key space (by minimizer)
	u32 slice_slots;	// S: slots of a slice's shared-memory image
	u32 m, w;		// minimizer length; m-mers of a window that are looked at: the central w of its K - m + 1
	u32 lo, wfull;		// first of them, K - m + 1
	u32 nmax;		// windows per record at most (32 for 1-word keys, else 64)
	u32 recw;		// u32 words per record (8, 12, 16)
	u32 slice_a;		// slice of the all-A k-mer (key 0), where the -n N-windows go (informational: the emit kernel gives N-windows that k-mer's minimizer value)
	u32 send_group;		// several GPUs, sending side: a chain holds this many consecutive slices and every record carries its slice in its last word (1: a chain is a slice)
	u32 tile_reads;		// reads per shared-memory tile of skm_emit_kernel
	u32 npos;		// m-mer positions per read at most (max_read_len - m + 1)
	u32 npad;		// row stride of the per-read arrays of skm_emit_kernel (a multiple of 4, >= npos)
};

static constexpr u32 SKM_HDR = 3;		// header words: ord low | ord high, n-1, flags, bases | slice
static constexpr u32 SKM_NFLAG = 0x80000000u;	// in the per-window slice array: window contains an N (-n)
static constexpr int EMIT_NT = 256;
static constexpr u32 EMIT_SEG = 16;	// windows per thread in the run detection of skm_emit_kernel

__host__ __device__ __forceinline__ u32 fmix32 (u32 h)
{
	h ^= h >> 16; h *= 0x85ebca6bu;
	h ^= h >> 13; h *= 0xc2b2ae35u;
	h ^= h >> 16;
	return h;
}
// hash of a canonical m-mer code, 31 bits; the window's minimizer value is the minimum of these.  Only the ORDER of
// the values matters (the slice comes from a full mix of the minimum, slice_of_min), and the high bits of an odd
// multiple depend on every bit of the code: one multiply is enough, and it runs once per base of every read.
__host__ __device__ __forceinline__ u32 mmer_hash (u32 code) { return ((code ^ 0x5bd1e995u) * 0x9E3779B1u) >> 1; }
#ifdef __CUDACC__
__device__ __forceinline__ u32 slice_of_min (u32 minval, u32 n_slices) { return __umulhi (fmix32 (minval + 0x9E3779B9u), n_slices); }
#endif
inline u32 slice_of_min_host (u32 minval, u32 n_slices) { return (u32) (((u64) fmix32 (minval + 0x9E3779B9u) * n_slices) >> 32); }

// window w of the tile -> read, offset, read length (the decode half of tile_chop)
template <bool NMODE>
__device__ __forceinline__ void tile_locate (const ReadTile<NMODE> &rt, int K, u32 w, u32 &r, u32 &j, u32 &len)
{
	if (rt.uniform)
	{
		r = w / rt.nwin_u;
		j = w - r * rt.nwin_u;
		len = rt.nwin_u + K - 1;
	}
	else
	{
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
		len = rt.prefix[r + 1] - rt.prefix[r] + K - 1;
	}
}

// the same with the division of the uniform case done by multiplication (mw = floor (2^32 / nwin_u) + 1, w < 65536)
template <bool NMODE>
__device__ __forceinline__ void skm_locate (const ReadTile<NMODE> &rt, int K, u32 mw, u32 w, u32 &r, u32 &j, u32 &len)
{
	if (rt.uniform)
	{
		r = __umulhi (w, mw);
		j = w - r * rt.nwin_u;
		len = rt.nwin_u + K - 1;
	}
	else
		tile_locate<NMODE> (rt, K, w, r, j, len);
}

// any N among bases [a, b) of a read (mask words: bit 31 of word 0 = base 0)
__device__ __forceinline__ bool mask_any (const u32 *mk, u32 a, u32 b)
{
	bool bad = false;
	for (u32 q = a >> 5; q <= (b - 1) >> 5; q++)
	{
		u32 x = mk[q];
		const u32 lo = q << 5;
		if (a > lo)
			x &= 0xFFFFFFFFu >> (a - lo);
		if (b < lo + 32)
			x &= 0xFFFFFFFFu << (lo + 32 - b);
		bad |= (x != 0);
	}
	return bad;
}
__device__ __forceinline__ bool mask_at (const u32 *mk, u32 b) { return (mk[b >> 5] >> (31 - (b & 31))) & 1u; }

// ------------------------------------------------------------------------------------------------
// reads -> super-k-mer records.  Record (recw u32 words, 16-byte aligned):
//   word 0  instance ordinal of the first window, low 32 bits
//   word 1  ordinal bits 32..39 | (n - 1) << 8 | has_left << 14 | n_run << 15 | n_bases << 16
//   word 2  0 (the multiplicity, once copies have been merged: skm_merge_kernel)
//   word 3.. bases, 16 per word, first base in the top bits (the layout of the read tile), starting
//           with the base before the first window if there is one (has_left): a miniature read on
//           which chop_window yields exactly the windows, keys and link bases of the original read.
//   n_run: n instances of key 0 without links (the -n N-windows, prlHashReads.c:193-196); no bases.
template <int W> struct SkmRec { static constexpr u32 WORDS = W == 1 ? 8u : (W == 2 ? 12u : 16u), NMAX = W == 1 ? 32u : 64u; };

template <int W, bool NMODE>
__global__ void __launch_bounds__ (EMIT_NT)
skm_emit_kernel (ReadBatch rb, SkmGeom g, SkmChains ch, unsigned long long *rec_cursor)
{
	constexpr u32 RECW = SkmRec<W>::WORDS, NMAX = SkmRec<W>::NMAX;
	extern __shared__ __align__(16) u32 smem[];
	__shared__ u32 warp_sums[EMIT_NT / 32];
	__shared__ u32 s_count, s_pool[2];
	pool_begin (ch, s_pool);
	const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int K = rb.K;
	ReadTile<NMODE> rt;
	tile_setup<NMODE> (rt, smem, rb);
	u32 *mhs = smem + tile_words (rb, NMODE);	// [tile_reads * npad + 16]: m-mer hashes, later descriptors (read | first window << 8 | (n - 1) << 24) of the records
	u32 *sls = mhs + (size_t) rb.tile_reads * g.npad + 16;	// [tile_reads * npad]: minimizer value of every window (31 bits | SKM_NFLAG); its slice is computed once per record
	const u32 maxwin = g.npos - g.wfull + 1;	// windows of the longest read
	const u32 gpr = (maxwin + 3) >> 2, spr = (maxwin + EMIT_SEG - 1) / EMIT_SEG;	// groups of 4 windows, segments of EMIT_SEG windows per read
	unsigned char *endb = reinterpret_cast<unsigned char *> (sls + (size_t) rb.tile_reads * g.npad);	// [tile_reads * 4 * spr]: per group of 4 windows, bit q: a run of windows ends at its window q
	for (u32 v = tid; v < rb.tile_reads * spr; v += EMIT_NT)
		reinterpret_cast<u32 *> (endb)[v] = 0u;	// (the groups beyond a row's last one are never written)
	const u32 m_npos = 0xFFFFFFFFu / g.npos + 1, m_gpr = 0xFFFFFFFFu / gpr + 1, m_spr = 0xFFFFFFFFu / spr + 1;	// x / d = (x * m) >> 32 for x < 65536
	const u64 n_tiles = (rb.n_reads + rb.tile_reads - 1) / rb.tile_reads;
	for (u64 t = blockIdx.x; t < n_tiles; t += gridDim.x)
	{
		tile_stage<NMODE, EMIT_NT> (rt, rb, t, warp_sums);
		if (tid == 0)
			s_count = 0;
		// ---- 1. hash of the canonical m-mer at every position of every read (m <= 15: 32-bit arithmetic)
		{
			const u32 msh = 32 - 2 * g.m;
			for (u32 x = tid; x < rt.nr * g.npos; x += EMIT_NT)
			{
				const u32 r = __umulhi (x, m_npos), i = x - r * g.npos;
				const u32 *rd = rt.tile + r * rt.sw + (i >> 4);
				const u32 f = __funnelshift_l (rd[1], rd[0], 2 * (i & 15)) >> msh;
				u32 c = __brev (f ^ 0xAAAAAAAAu);	// complement (x ^ 2 per base, inc/def.h:42), then reverse the 2-bit groups
				c = (((c >> 1) & 0x55555555u) | ((c & 0x55555555u) << 1)) >> msh;
				mhs[r * g.npad + i] = mmer_hash (min (f, c));
			}
		}
		__syncthreads ();
		// ---- 2. minimizer value of every window (its slice follows from it): the smallest hash among the CENTRAL w of the
		// window's m-mers, positions lo .. lo + w - 1 (a set that is the same for a k-mer and its reverse
		// complement).  For K <= 31 that is every m-mer of the window; for long k-mers it keeps a minimizer from
		// gathering the windows of 50 or 110 consecutive positions in one slice (K = 63: 11 % of the slices overflowed).
		// One thread per four consecutive windows j0 .. j0+3: they share the hashes j0+3 .. j0+w-1, so the
		// sliding minimum costs ~(w + 9) / 4 comparisons per window instead of w; 16-byte loads keep the
		// stride-4 access free of bank conflicts.
		for (u32 x = tid; x < rt.nr * gpr; x += EMIT_NT)
		{
			const u32 r = gpr == 1 ? x : __umulhi (x, m_gpr), j0 = 4 * (x - r * gpr);	// (the multiplier of 1 does not fit 32 bits)
			const u32 nwin = rt.uniform ? rt.nwin_u : rt.prefix[r + 1] - rt.prefix[r];
			if (j0 >= nwin)
			{
				endb[(r * spr << 2) + (j0 >> 2)] = 0;	// (the row's read in the previous tile may have been longer)
				continue;
			}
			const uint4 *mh = reinterpret_cast<const uint4 *> (mhs + r * g.npad + j0 + g.lo);	// (lo is a multiple of 4)
			const uint4 c0 = mh[0];
			const u32 a2 = c0.z, a1 = min (c0.y, a2), a0 = min (c0.x, a1);	// hashes 0..2 belong to the first windows only
			// hash 3 and hashes 4 .. w-1 apart: the FIFTH window (the next group's first; computed here only to
			// tell whether a run of windows ends at this group's fourth) does without hash 3
			u32 c3 = g.w > 3 ? c0.w : 0xFFFFFFFFu, common4 = 0xFFFFFFFFu;
			u32 b0 = 0xFFFFFFFFu, b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu, b3 = 0xFFFFFFFFu;	// running minima of hashes w, w..w+1, w..w+2, w..w+3
			if (g.w == 3)
				b0 = b1 = b2 = b3 = c0.w;
			const u32 nv = (g.w + 7) >> 2;	// 16-byte chunks that hold hashes 0 .. w+3
			for (u32 k = 1; k < nv; k++)
			{
				const uint4 c = mh[k];
				const u32 e = 4 * k;
				if (e + 3 < g.w)
					common4 = min (common4, min (min (c.x, c.y), min (c.z, c.w)));
				else
				{	// the chunk reaches past hash w-1
					const u32 v[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
					for (u32 q = 0; q < 4; q++)
					{
						const u32 idx = e + q;
						if (idx < g.w)
							common4 = min (common4, v[q]);
						else
						{
							if (idx == g.w)
								b0 = v[q];
							if (idx <= g.w + 1)
								b1 = min (b1, v[q]);
							if (idx <= g.w + 2)
								b2 = min (b2, v[q]);
							if (idx <= g.w + 3)
								b3 = min (b3, v[q]);
						}
					}
				}
			}
			b1 = min (b1, b0);
			b2 = min (b2, b1);
			b3 = min (b3, b2);
			const u32 common = min (c3, common4);
			u32 mv[5] = { min (common, a0), min (min (common, a1), b0), min (min (common, a2), b1), min (common, b2), min (common4, b3) };
			if constexpr (NMODE)
			{
#pragma unroll
				for (u32 q = 0; q < 5; q++)
					if (j0 + q < nwin && mask_any (rt.mtile + r * rt.mw, j0 + q, j0 + q + K))
						mv[q] = mmer_hash (0) | SKM_NFLAG;	// with the all-A k-mer (key 0), whose every m-mer has code 0
			}
			u32 *sl = sls + r * g.npad + j0;
			u32 eb = 0;
#pragma unroll
			for (u32 q = 0; q < 4; q++)
				if (j0 + q < nwin)
				{
					sl[q] = mv[q];
					eb |= (j0 + q + 1 >= nwin || mv[q + 1] != mv[q] ? 1u : 0u) << q;
				}
			endb[(r * spr << 2) + (j0 >> 2)] = (unsigned char) eb;
		}
		__syncthreads ();
		// ---- 3. runs of windows with the same minimizer value -> record descriptors in mhs[], which is free now.
		// One thread per segment of EMIT_SEG windows of a read walks them in order and reports the runs
		// that END in its segment (a run that started before the segment is traced back to its start).
		for (u32 x = tid; x < rt.nr * spr; x += EMIT_NT)
		{
			const u32 r = spr == 1 ? x : __umulhi (x, m_spr), js = EMIT_SEG * (x - r * spr);
			const u32 nwin = rt.uniform ? rt.nwin_u : rt.prefix[r + 1] - rt.prefix[r];
			if (js >= nwin)
				continue;
			const u32 *sl = sls + r * g.npad;
			const u32 *ew = reinterpret_cast<const u32 *> (endb) + r * spr;	// a segment's four groups: one word, byte b bit q = window js + 4 b + q
			u32 ends = ew[js / EMIT_SEG], start = 0;
			// the run that is open at the segment's first window started behind the last end before the segment
			for (int w = (int) (js / EMIT_SEG) - 1; w >= 0; w--)
				if (const u32 e2 = ew[w])
				{
					const u32 p = 31 - __clz (e2);
					start = EMIT_SEG * (u32) w + ((p >> 3) << 2) + (p & 7u) + 1;
					break;
				}
			while (ends)
			{
				const u32 p = (u32) __ffs (ends) - 1;
				const u32 j = js + ((p >> 3) << 2) + (p & 7u);	// windows start .. j are a run
				ends &= ends - 1;
				if (rb.owner_ranks <= 1 || slice_of_min (sl[j] & ~SKM_NFLAG, g.n_slices) % rb.owner_ranks == rb.owner_rank)
				{
					const u32 n = j - start + 1;
					const u32 nrec = (n + NMAX - 1) / NMAX;
					const u32 pos = atomicAdd (&s_count, nrec);
					mhs[pos] = r | (start << 8) | ((min (NMAX, n) - 1) << 24);
					for (u32 c = 1; c < nrec; c++)	// runs longer than a record holds: rare
						mhs[pos + c] = r | ((start + c * NMAX) << 8) | ((min (NMAX, n - c * NMAX) - 1) << 24);
				}
				start = j + 1;
			}
		}
		__syncthreads ();
		if (tid == 0)
		{
			if (s_count)
				atomicAdd (rec_cursor, (unsigned long long) s_count);	// records made (also by tiles that find the pool exhausted)
			pool_refill (ch, s_pool);
		}
		__syncthreads ();
		const u32 n_out = s_count;
		if (s_pool[1] == 0)
		{	// the block pool is exhausted: the host enlarges it and emits the whole read log again (sdtgpu.cu)
			__syncthreads ();
			continue;
		}
		// ---- 4. one thread per record: the record goes to the end of its slice's chain
		// (the position of a thread's NEXT record is drawn before this one is assembled: the atomic's round trip
		// to the L2 is covered by the assembly and the stores)
		u32 d_nx = 0, s_nx = 0;
		u64 tk_nx = 0;
		if (tid < n_out)
		{
			d_nx = mhs[tid];
			s_nx = sls[(d_nx & 0xFFu) * g.npad + ((d_nx >> 8) & 0xFFFFu)];
			s_nx = slice_of_min (s_nx & ~SKM_NFLAG, g.n_slices) | (s_nx & SKM_NFLAG);
			tk_nx = chain_ticket (ch, (s_nx & ~SKM_NFLAG) / g.send_group);
		}
		for (u32 rid = tid; rid < n_out; rid += EMIT_NT)
		{
			const u32 d = d_nx, s = s_nx;
			const u64 tk = tk_nx;
			const bool more = rid + EMIT_NT < n_out;
			if (more)
			{
				d_nx = mhs[rid + EMIT_NT];
				s_nx = sls[(d_nx & 0xFFu) * g.npad + ((d_nx >> 8) & 0xFFFFu)];
				s_nx = slice_of_min (s_nx & ~SKM_NFLAG, g.n_slices) | (s_nx & SKM_NFLAG);
				tk_nx = chain_ticket (ch, (s_nx & ~SKM_NFLAG) / g.send_group);
			}
			const u32 r = d & 0xFFu, j0 = (d >> 8) & 0xFFFFu, n = (d >> 24) + 1;
			const u32 len = (rt.uniform ? rt.nwin_u : rt.prefix[r + 1] - rt.prefix[r]) + K - 1;
			const u32 *rd = rt.tile + r * rt.sw;
			u32 *rec = chain_place (ch, (s & ~SKM_NFLAG) / g.send_group, tk, s_pool, more ? &tk_nx : nullptr, (s_nx & ~SKM_NFLAG) / g.send_group);
			if (!rec)
				continue;
			const u64 ord = (rb.first_read_ordinal + rt.r0 + r) * rb.maxwin + j0;
			const bool nrun = NMODE && (s & SKM_NFLAG);
			u32 has_left = j0 > 0, has_right = j0 + n - 1 + K < len;
			if constexpr (NMODE)
			{
				const u32 *mk = rt.mtile + r * rt.mw;
				if (has_left && mask_at (mk, j0 - 1))
					has_left = 0;
				if (has_right && mask_at (mk, j0 + n - 1 + K))
					has_right = 0;
			}
			const u32 nb = nrun ? 0 : has_left + K + n - 1 + has_right;
			const u32 first = j0 - has_left;
			u32 wd[4];
			wd[0] = (u32) ord;
			wd[1] = (u32) (ord >> 32) | ((n - 1) << 8) | (has_left << 14) | ((nrun ? 1u : 0u) << 15) | (nb << 16);
			wd[2] = 0;
#pragma unroll
			for (u32 q = 0; q < RECW - SKM_HDR; q++)
			{
				// (no branch on the record's length: the words past its bases are read from whatever follows in the
				// tile — shared memory of this CTA in any case — and masked away, so the warp stays converged)
				const u32 b = first + 16 * q, wq = b >> 4, sh = 2 * (b & 15);
				u32 v = __funnelshift_l (rd[wq + 1], rd[wq], sh);
				const u32 left = nb > 16 * q ? nb - 16 * q : 0;	// bases of the record from this word on
				v = left >= 16 ? v : left ? v & (0xFFFFFFFFu << (32 - 2 * left)) : 0;	// nothing of the read beyond the record's bases
				const u32 o = SKM_HDR + q;
				if (q == RECW - SKM_HDR - 1 && g.send_group > 1)
					v = s & ~SKM_NFLAG;	// (the last word is never used by bases: the slice travels there)
				wd[o & 3] = v;
				if ((o & 3) == 3)
					*reinterpret_cast<uint4 *> (rec + (o & ~3u)) = make_uint4 (wd[0], wd[1], wd[2], wd[3]);
			}
		}
		__syncthreads ();	// tile, descriptors and slices are overwritten by the next iteration
	}
	pool_end (ch, s_pool);
}

// ------------------------------------------------------------------------------------------------
// The slice's table image in shared memory.  Shared-memory atomics are the scarce resource (2 cycles
// per lane, 20x a load), so an instance costs exactly ONE: every slot has a 5 x 5 matrix of 16-bit
// cells indexed by (left, right), 4 = "no neighbour".  The reference's update_kmer (newhash.c:71-96)
// touches count, one left and one right counter per instance; all three are sums over that matrix:
//     count = sum of all cells + extra (mod 2^32),  L[b] = min (63, row b),  R[b] = min (63, column b).
// A cell that gets near 16 bits stops counting and the slot's 32-bit `extra` takes over (it only
// feeds count: a saturated cell already pins its row and column at 63).
static constexpr int CELL_WORDS = 13;	// 25 16-bit cells, two per word
static constexpr u32 SKM_MAX_TRIES = 128;
static constexpr u32 CELL_STOP = 0xF000u;	// + one in-flight increment per thread of the CTA stays below 2^16

// two cheap 32-bit hashes of a key: a -> home slot, b -> probe step, retry filter, sub-slice
template <int W> __device__ __forceinline__ void skm_hash2 (const Key<W> &k, u32 &a, u32 &b)
{
	u32 x = 0;
#pragma unroll
	for (int q = 0; q < W; q++)
	{
		x = (x ^ (u32) k.w[q]) * 0x9E3779B1u;
		x = (x ^ (u32) (k.w[q] >> 32)) * 0x85EBCA6Bu + (x >> 17);
	}
	x ^= x >> 15; x *= 0xC2B2AE35u;
	x ^= x >> 13;
	a = x;
	b = x * 0x27D4EB2Fu;
	b ^= b >> 15;
}

// a work item of the build: records r0 .. r0 + nrec of the merged runs (all instances of their k-mers), wsum windows;
// of these the keys with sub-hash % R == r (a piece of an item that overflowed an image); R == 0: do not even try
struct SkmWork { u64 r0; u32 nrec, wsum, r, R; };

// Rolling state of one record: what nextKmer / reverseComplement (kmer.c:209, 653) compute per base,
// kept incrementally — the forward k-mer takes the next base at the bottom, its reverse complement
// takes the complement at the top — so a window costs a handful of shifts instead of an extraction.
template <int W> struct SkmRoll
{
	static constexpr int PW = W == 1 ? 1 : 2;	// 64-bit words of bases still to come (NMAX bases at most)
	Key<W> f, rc;
	u64 pend[PW];	// the bases after the current window, first one in the top bits
	u64 ord;	// instance ordinal of the current window
	u32 left;	// base before the current window (4: none)
	u32 n, t;	// windows of the record, current window
	u32 has_right;	// the last window has a base after it
	u32 add;	// instances per step: the record's multiplicity (skm_dedupe_kernel); an N-run applies all of its windows at once
};

// bits [2 * b, 2 * b + 64) of the record's bases (words rd[0 .. last], first base in the top bits);
// words past `last` hold nothing a window can ask for and are not read
__device__ __forceinline__ u64 bases64 (const u32 *rd, u32 b, u32 last)
{
	const u32 q = b >> 4, r = 2 * (b & 15);
	const u32 w0 = rd[min (q, last)], w1 = rd[min (q + 1, last)], w2 = rd[min (q + 2, last)];
	return ((u64) __funnelshift_l (w1, w0, r) << 32) | __funnelshift_l (w2, w1, r);
}

// rolling state at window tw of a record
template <int W, bool FAST = true>
__device__ __forceinline__ void skm_roll_init (SkmRoll<W> &s, const u32 *rec, int K, u32 tw)
{
	constexpr u32 LAST = SkmRec<W>::WORDS - SKM_HDR - 1;
	const uint4 hd = __ldg (reinterpret_cast<const uint4 *> (rec));	// ord low | header | slice | first bases
	const u32 h1 = hd.y;
	const u32 n = ((h1 >> 8) & 63u) + 1, has_left = (h1 >> 14) & 1u, nrun = (h1 >> 15) & 1u, nb = h1 >> 16;
	s.ord = ((u64) hd.x | ((u64) (h1 & 0xFFu) << 32)) + tw;
	s.t = tw;
	if (nrun)
	{	// n instances of key 0 without links: they only feed count (and the ordinal); one step (tw = 0)
#pragma unroll
		for (int q = 0; q < W; q++)
		{
			s.f.w[q] = 0;
			s.rc.w[q] = ~0ull;
		}
#pragma unroll
		for (int q = 0; q < SkmRoll<W>::PW; q++)
			s.pend[q] = 0;
		s.left = 4;
		s.n = 1;
		s.has_right = 0;
		s.add = n * hd.z;
		return;
	}
	if constexpr (W == 1 && FAST)
		if (tw == 0)
		{	// start of a record (the common re-seat): the record's 160 bits of bases, two 16-byte loads
			const uint4 h2 = __ldg (reinterpret_cast<const uint4 *> (rec) + 1);
			u64 hi = ((u64) hd.w << 32) | h2.x, mid = ((u64) h2.y << 32) | h2.z;
			if (has_left)
			{
				hi = (hi << 2) | (mid >> 62);
				mid = (mid << 2) | ((u64) h2.w >> 30);
			}
			const int sh = 64 - 2 * K;	// 2 <= sh
			s.f.w[0] = hi >> sh;
			s.rc.w[0] = revcomp64 (s.f.w[0]) >> sh;
			s.pend[0] = (hi << (64 - sh)) | (mid >> sh);
			s.left = has_left ? hd.w >> 30 : 4u;
			s.n = n;
			s.has_right = nb - has_left - (u32) K - (n - 1);
			s.add = hd.z;
			return;
		}
	const u32 *rd = rec + SKM_HDR;
	const u32 j = has_left + tw;	// first base of the window
	const u32 p0 = j + (u32) K;	// first base after it
	extract_fwd<W> (rd, (int) p0, K, s.f);
	revcomp<W> (s.f, K, s.rc);
	s.left = j ? base_at (rd, (int) j - 1) : 4u;
	s.pend[0] = bases64 (rd, p0, LAST);
	if constexpr (W > 1)
		s.pend[1] = n - tw > 32 ? bases64 (rd, p0 + 32, LAST) : 0ull;	// only long records reach into the second word
	s.n = n;
	s.has_right = nb - has_left - (u32) K - (n - 1);
	s.add = hd.z;
}

// the window's canonical key and its links in the stored orientation (chopKmer4read, prlHashReads.c:215-230, 275-308)
template <int W>
__device__ __forceinline__ void skm_roll_window (const SkmRoll<W> &s, Key<W> &key, u32 &left, u32 &right)
{
	const u32 nb = (u32) (s.pend[0] >> 62);
	const u32 next = (s.t + 1 == s.n && !s.has_right) ? 4u : nb;
	if (key_less<W> (s.f, s.rc))
	{
		key = s.f;
		left = s.left;
		right = next;
	}
	else
	{
		key = s.rc;
		left = next < 4 ? (next ^ 2u) : 4u;
		right = s.left < 4 ? (s.left ^ 2u) : 4u;
	}
}

// one base forward.  top = 2 * (K - 1): bit position of the k-mer's first base
template <int W>
__device__ __forceinline__ void skm_roll_step (SkmRoll<W> &s, const Key<W> &mask, int top)
{
	const u64 nb = s.pend[0] >> 62;
	const int tw = W - 1 - (top >> 6), ts = top & 63;	// word and shift of the first base
#pragma unroll
	for (int q = 0; q < W; q++)
		if (q == tw)
			s.left = (u32) (s.f.w[q] >> ts) & 3u;
#pragma unroll
	for (int q = 0; q < W; q++)
		s.f.w[q] = ((s.f.w[q] << 2) | (q + 1 < W ? s.f.w[q + 1] >> 62 : nb)) & mask.w[q];
#pragma unroll
	for (int q = W - 1; q >= 0; q--)
	{
		u64 v = s.rc.w[q] >> 2;
		if (q > 0)
			v |= s.rc.w[q - 1] << 62;
		if (q == tw)
			v |= (nb ^ 2ull) << ts;
		s.rc.w[q] = v;
	}
	if constexpr (W == 1)
		s.pend[0] <<= 2;
	else
	{
		s.pend[0] = (s.pend[0] << 2) | (s.pend[1] >> 62);
		s.pend[1] <<= 2;
	}
	s.t++;
	s.ord++;
}

}	// namespace sdt
