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

namespace sdt {

struct SkmGeom
{
	u32 n_slices;		// slices of the key space (by minimizer)
	u32 slice_slots;	// S: slots of a slice's shared-memory image
	u32 m, w;		// minimizer length, m-mers per window (K - m + 1)
	u32 nmax;		// windows per record at most (32 for 1-word keys, else 64)
	u32 recw;		// u32 words per record (8, 12, 16)
	u32 slice_a;		// slice of the all-A k-mer (key 0): where the -n N-windows go
	u32 tile_reads;		// reads per shared-memory tile of skm_emit_kernel
	u32 npos;		// m-mer positions per read at most (max_read_len - m + 1)
	u32 npad;		// row stride of the per-read arrays of skm_emit_kernel (a multiple of 4, >= npos)
	u32 build_nt;		// threads per CTA of skm_build_kernel: 1024 (one CTA per SM) or 512 (two)
};

static constexpr u32 SKM_HDR = 3;		// header words: ord low | ord high, n-1, flags, bases | slice
static constexpr u32 SKM_NFLAG = 0x80000000u;	// in the per-window slice array: window contains an N (-n)
static constexpr int EMIT_NT = 256;
static constexpr u32 EMIT_SEG = 16;	// windows per thread in the run detection of skm_emit_kernel
static constexpr int SCAT_NT = 256;

__host__ __device__ __forceinline__ u32 fmix32 (u32 h)
{
	h ^= h >> 16; h *= 0x85ebca6bu;
	h ^= h >> 13; h *= 0xc2b2ae35u;
	h ^= h >> 16;
	return h;
}
// hash of a canonical m-mer code; the window's minimizer value is the minimum of these
__host__ __device__ __forceinline__ u32 mmer_hash (u32 code) { return fmix32 (code ^ 0x5bd1e995u); }
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
//   word 2  slice
//   word 3.. bases, 16 per word, first base in the top bits (the layout of the read tile), starting
//           with the base before the first window if there is one (has_left): a miniature read on
//           which chop_window yields exactly the windows, keys and link bases of the original read.
//   n_run: n instances of key 0 without links (the -n N-windows, prlHashReads.c:193-196); no bases.
template <int W> struct SkmRec { static constexpr u32 WORDS = W == 1 ? 8u : (W == 2 ? 12u : 16u), NMAX = W == 1 ? 32u : 64u; };

template <int W, bool NMODE>
__global__ void __launch_bounds__ (EMIT_NT)
skm_emit_kernel (ReadBatch rb, SkmGeom g, u32 *hist, u32 *rec0, u64 rec_cap, unsigned long long *rec_cursor, Counters *ctr)
{
	constexpr u32 RECW = SkmRec<W>::WORDS, NMAX = SkmRec<W>::NMAX;
	extern __shared__ __align__(16) u32 smem[];
	__shared__ u32 warp_sums[EMIT_NT / 32];
	__shared__ u32 s_count;
	__shared__ unsigned long long s_base;
	const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int K = rb.K;
	ReadTile<NMODE> rt;
	tile_setup<NMODE> (rt, smem, rb);
	u32 *mhs = smem + tile_words (rb, NMODE);	// [tile_reads * npad + 16]: m-mer hashes, later descriptors (read | first window << 8 | (n - 1) << 24) of the records
	u32 *sls = mhs + (size_t) rb.tile_reads * g.npad + 16;	// [tile_reads * npad]: slice of every window (| SKM_NFLAG)
	const u32 maxwin = g.npos - g.w + 1;	// windows of the longest read
	const u32 gpr = (maxwin + 3) >> 2, spr = (maxwin + EMIT_SEG - 1) / EMIT_SEG;	// groups of 4 windows, segments of EMIT_SEG windows per read
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
		// ---- 2. slice of every window = slice of its minimizer value (the smallest of its w m-mer hashes).
		// One thread per four consecutive windows j0 .. j0+3: they share the hashes j0+3 .. j0+w-1, so the
		// sliding minimum costs ~(w + 9) / 4 comparisons per window instead of w; 16-byte loads keep the
		// stride-4 access free of bank conflicts.
		for (u32 x = tid; x < rt.nr * gpr; x += EMIT_NT)
		{
			const u32 r = gpr == 1 ? x : __umulhi (x, m_gpr), j0 = 4 * (x - r * gpr);	// (the multiplier of 1 does not fit 32 bits)
			const u32 nwin = rt.uniform ? rt.nwin_u : rt.prefix[r + 1] - rt.prefix[r];
			if (j0 >= nwin)
				continue;
			const uint4 *mh = reinterpret_cast<const uint4 *> (mhs + r * g.npad + j0);
			const uint4 c0 = mh[0];
			const u32 a2 = c0.z, a1 = min (c0.y, a2), a0 = min (c0.x, a1);	// hashes 0..2 belong to the first windows only
			u32 common = g.w > 3 ? c0.w : 0xFFFFFFFFu;	// hashes 3 .. w-1: in all four windows
			u32 b0 = 0xFFFFFFFFu, b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;	// running minima of hashes w, w..w+1, w..w+2
			if (g.w == 3)
				b0 = b1 = b2 = c0.w;
			const u32 nv = (g.w + 6) >> 2;	// 16-byte chunks that hold hashes 0 .. w+2
			for (u32 k = 1; k < nv; k++)
			{
				const uint4 c = mh[k];
				const u32 e = 4 * k;
				if (e + 3 < g.w)
					common = min (common, min (min (c.x, c.y), min (c.z, c.w)));
				else
				{	// the chunk reaches past hash w-1
					const u32 v[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
					for (u32 q = 0; q < 4; q++)
					{
						const u32 idx = e + q;
						if (idx < g.w)
							common = min (common, v[q]);
						else
						{
							if (idx == g.w)
								b0 = v[q];
							if (idx <= g.w + 1)
								b1 = min (b1, v[q]);
							if (idx <= g.w + 2)
								b2 = min (b2, v[q]);
						}
					}
				}
			}
			b1 = min (b1, b0);
			b2 = min (b2, b1);
			const u32 mv[4] = { min (common, a0), min (min (common, a1), b0), min (min (common, a2), b1), min (common, b2) };
			u32 *sl = sls + r * g.npad + j0;
#pragma unroll
			for (u32 q = 0; q < 4; q++)
				if (j0 + q < nwin)
				{
					u32 v = slice_of_min (mv[q], g.n_slices);
					if constexpr (NMODE)
						if (mask_any (rt.mtile + r * rt.mw, j0 + q, j0 + q + K))
							v = g.slice_a | SKM_NFLAG;
					sl[q] = v;
				}
		}
		__syncthreads ();
		// ---- 3. runs of windows with the same slice -> record descriptors in mhs[], which is free now.
		// One thread per segment of EMIT_SEG windows of a read walks them in order and reports the runs
		// that END in its segment (a run that started before the segment is traced back to its start).
		for (u32 x = tid; x < rt.nr * spr; x += EMIT_NT)
		{
			const u32 r = spr == 1 ? x : __umulhi (x, m_spr), js = EMIT_SEG * (x - r * spr);
			const u32 nwin = rt.uniform ? rt.nwin_u : rt.prefix[r + 1] - rt.prefix[r];
			if (js >= nwin)
				continue;
			const u32 je = min (js + EMIT_SEG, nwin);
			const u32 *sl = sls + r * g.npad;
			u32 s = sl[js], start = js;
			while (start > 0 && sl[start - 1] == s)
				start--;
			// first mark where runs end (a tight, convergent loop), then report them: reporting inside the
			// walk made every step of the warp pay for the few lanes that had a run to report
			u32 ends = 0;
			for (u32 j = js; j < je; j++)
			{
				const u32 nxt = j + 1 < nwin ? sl[j + 1] : 0xFFFFFFFFu;	// no slice has this number (bit 31 is the N flag, 30 bits of slice)
				ends |= (nxt != s ? 1u : 0u) << (j - js);
				s = nxt;
			}
			while (ends)
			{
				const u32 j = js + (u32) __ffs (ends) - 1;	// windows start .. j are a run
				ends &= ends - 1;
				const u32 rs = sl[j];
				if (rb.owner_ranks <= 1 || (rs & ~SKM_NFLAG) % rb.owner_ranks == rb.owner_rank)
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
			s_base = s_count ? atomicAdd (rec_cursor, (unsigned long long) s_count) : 0ull;
		__syncthreads ();
		const u64 base = s_base;
		const u32 n_out = s_count;
		if (base + n_out > rec_cap)
		{	// the host re-emits the whole read log into a larger area (sdtgpu.cu)
			if (tid == 0)
				atomicOr (reinterpret_cast<unsigned long long *> (&ctr->overflow), 2ull);
			__syncthreads ();
			continue;
		}
		// ---- 4. one thread per record: consecutive threads write consecutive records
		for (u32 rid = tid; rid < n_out; rid += EMIT_NT)
		{
			const u32 d = mhs[rid];
			const u32 r = d & 0xFFu, j0 = (d >> 8) & 0xFFFFu, n = (d >> 24) + 1;
			const u32 len = (rt.uniform ? rt.nwin_u : rt.prefix[r + 1] - rt.prefix[r]) + K - 1;
			const u32 s = sls[r * g.npad + j0];
			const u32 *rd = rt.tile + r * rt.sw;
			u32 *rec = rec0 + (base + rid) * RECW;
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
			wd[2] = s & ~SKM_NFLAG;
			const u32 nbw = (nb + 15) >> 4;
#pragma unroll
			for (u32 q = 0; q < RECW - SKM_HDR; q++)
			{
				u32 v = 0;
				if (q < nbw)
				{
					const u32 b = first + 16 * q, wq = b >> 4, sh = 2 * (b & 15);
					v = __funnelshift_l (rd[wq + 1], rd[wq], sh);
					if (q == nbw - 1 && (nb & 15))
						v &= 0xFFFFFFFFu << (32 - 2 * (nb & 15));	// nothing of the read beyond the record's bases
				}
				const u32 o = SKM_HDR + q;
				wd[o & 3] = v;
				if ((o & 3) == 3)
					*reinterpret_cast<uint4 *> (rec + (o & ~3u)) = make_uint4 (wd[0], wd[1], wd[2], wd[3]);
			}
			atomicAdd (hist + (s & ~SKM_NFLAG), 1u);	// RED
		}
		__syncthreads ();	// tile, descriptors and slices are overwritten by the next iteration
	}
}

// every record to its slice's run: cur[p] starts at off[p].  The kernel is bound by the rate at which
// the memory system takes requests to cold lines (profiles/r1_random_access_findings.md), so a record
// goes out in as few stores as its size allows: one 32-byte store for 1-word keys, two for 4-word keys.
__device__ __forceinline__ void ld256_stream (const void *p, u64 &a, u64 &b, u64 &c, u64 &d)
{
	asm volatile ("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

template <int RECW>
__global__ void __launch_bounds__ (SCAT_NT)
skm_scatter_kernel (const u32 *rec0, const unsigned long long *rec_count, unsigned long long *cur, u32 *rec2)
{
	const u64 n = *rec_count;
	for (u64 i = blockIdx.x * (u64) SCAT_NT + threadIdx.x; i < n; i += (u64) gridDim.x * SCAT_NT)
	{
		const u32 *src = rec0 + i * RECW;
		if constexpr (RECW % 8 == 0)
		{
			u64 a, b, c, d;
			ld256_stream (src, a, b, c, d);
			const u64 pos = atomicAdd (cur + (u32) b, 1ull);	// word 2: the slice
			u32 *dst = rec2 + pos * RECW;
			st256 (dst, a, b, c, d);
#pragma unroll
			for (int q = 1; q < RECW / 8; q++)
			{
				ld256_stream (src + 8 * q, a, b, c, d);
				st256 (dst + 8 * q, a, b, c, d);
			}
		}
		else
		{
			const uint4 *s4 = reinterpret_cast<const uint4 *> (src);
			const uint4 h = ldg_stream (s4);
			const u64 pos = atomicAdd (cur + h.z, 1ull);
			uint4 *dst = reinterpret_cast<uint4 *> (rec2 + pos * RECW);
			dst[0] = h;
#pragma unroll
			for (int q = 1; q < RECW / 4; q++)
				dst[q] = ldg_stream (s4 + q);
		}
	}
}

// super-k-mer exchange, receiving side: records that arrived from all ranks carry global slice
// numbers; this rank owns slices [lo, lo + n_local).  Re-base them and count records per slice.
__global__ void __launch_bounds__ (SCAT_NT)
skm_recount_kernel (u32 *rec, u64 n, u32 recw, u32 lo, u32 n_local, u32 *hist, Counters *ctr)
{
	for (u64 i = blockIdx.x * (u64) SCAT_NT + threadIdx.x; i < n; i += (u64) gridDim.x * SCAT_NT)
	{
		u32 *w2 = rec + i * recw + 2;
		const u32 s = *w2 - lo;
		if (s >= n_local)
		{	// not this rank's: the exchange went wrong
			atomicOr (reinterpret_cast<unsigned long long *> (&ctr->overflow), 16ull);
			continue;
		}
		*w2 = s;
		atomicAdd (hist + s, 1u);	// RED
	}
}

// ------------------------------------------------------------------------------------------------
// Copies of the same super-k-mer collapse into one record before the build.  At the coverage of a
// transcriptome most records of a slice are byte-identical copies (every error-free read that spans
// a super-k-mer emits the same bases, neighbours and window count; only the ordinal differs), and a
// window costs the build ~400 instructions while finding a copy costs ~100 per record.  A surviving
// record carries its multiplicity in word 2 (the slice number is no longer needed once the record
// sits in its slice's run) and the smallest ordinal of its copies: window t of every copy is the
// same (key, left, right) instance with ordinal ord0 + t, so count and link counters take the
// multiplicity (update_kmer, newhash.c:71-96, is a sum) and the node's first ordinal the minimum.
//
// One CTA per slice, DD_CHUNK records at a time staged in shared memory; a table of record indices
// keyed by the record's content finds the copies; survivors go back to the front of the run.
// end[slice] = one past the last surviving record.
static constexpr int DD_NT = 512;	// 2 CTAs of 80-96 KB per SM: 32 warps to hide the staging loads (256 threads: 11.2 ms on C2)
template <int W> struct DedupeCfg { static constexpr u32 CHUNK = W == 1 ? 2048u : 1024u, TABLE = 2 * CHUNK; };
template <int W> __host__ __device__ inline size_t skm_dedupe_smem () { return (size_t) DedupeCfg<W>::CHUNK * SkmRec<W>::WORDS * 4 + (size_t) DedupeCfg<W>::TABLE * 4; }

// HAS_MULT: word 2 of the incoming records already is a multiplicity (the sub-records of skm_resplit_kernel)
template <int W, bool HAS_MULT>
__global__ void __launch_bounds__ (DD_NT)
skm_dedupe_kernel (u32 *rec2, const u64 *off, u32 n_slices, unsigned long long *end, unsigned long long *n_kept)
{
	constexpr u32 RECW = SkmRec<W>::WORDS, CH = DedupeCfg<W>::CHUNK, TS = DedupeCfg<W>::TABLE, VEC = RECW / 4;
	extern __shared__ __align__(16) u32 smem[];
	__shared__ u32 s_out;
	u32 *st = smem;			// [CH * RECW]: the chunk's records
	u32 *tab = smem + CH * RECW;	// [TS]: record index + 1, 0 = free
	const u32 tid = threadIdx.x, lane = tid & 31;
	u64 kept_total = 0;	// thread 0
	for (u32 sl = blockIdx.x; sl < n_slices; sl += gridDim.x)
	{
		const u64 r0 = off[sl], r1 = off[sl + 1];
		u64 out = r0;	// next free position of the run
		for (u64 c0 = r0; c0 < r1; c0 += CH)
		{
			const u32 nrec = (u32) min ((u64) CH, r1 - c0);
			{	// stage (coalesced 16-byte loads); word 2 becomes the multiplicity
				const uint4 *src = reinterpret_cast<const uint4 *> (rec2 + c0 * RECW);
				uint4 *dst = reinterpret_cast<uint4 *> (st);
				for (u32 v = tid; v < nrec * VEC; v += DD_NT)
				{
					uint4 x = ldg_stream (src + v);
					if (!HAS_MULT && v % VEC == 0)
						x.z = 1u;
					dst[v] = x;
				}
				for (u32 v = tid; v < TS; v += DD_NT)
					tab[v] = 0u;
				if (tid == 0)
					s_out = 0;
			}
			__syncthreads ();
			for (u32 i = tid; i < nrec; i += DD_NT)
			{
				u32 *me = st + i * RECW;
				const u32 h1 = me[1];
				if ((h1 >> 15) & 1u)
					continue;	// an N-run stays as it is
				u32 hsh = h1 & ~0xFFu;
#pragma unroll
				for (u32 q = SKM_HDR; q < RECW; q++)
					hsh = (hsh ^ me[q]) * 0x9E3779B1u + (hsh >> 15);
				u32 slot = fmix32 (hsh) & (TS - 1);
				for (;;)
				{
					u32 e = tab[slot];
					if (e == 0u)
						e = atomicCAS (tab + slot, 0u, i + 1);
					if (e == 0u)
						break;	// first of its kind
					u32 *rep = st + (e - 1) * RECW;
					bool same = ((*reinterpret_cast<volatile u32 *> (rep + 1) ^ h1) & ~0xFFu) == 0u;	// the low 8 bits are ordinal bits and change
#pragma unroll
					for (u32 q = SKM_HDR; q < RECW; q++)
						same &= rep[q] == me[q];
					if (same)
					{	// words 0-1 as one 64-bit number: the header bits above the ordinal are equal, so the minimum is the ordinal's
						atomicAdd (rep + 2, HAS_MULT ? me[2] : 1u);
						const u64 mine = *reinterpret_cast<const u64 *> (me);
						if (mine < *reinterpret_cast<volatile u64 *> (rep))
							atomicMin (reinterpret_cast<unsigned long long *> (rep), mine);
						me[2] = 0u;	// dropped
						break;
					}
					slot = (slot + 1) & (TS - 1);
				}
			}
			__syncthreads ();
			// survivors back to the run, in any order: a warp reserves its share with one atomic
			for (u32 b = 0; b < nrec; b += DD_NT)
			{
				const u32 i = b + tid;
				const bool keep = i < nrec && st[i * RECW + 2] != 0u;
				const u32 bal = __ballot_sync (0xFFFFFFFFu, keep);
				u32 base = 0;
				if (lane == 0 && bal)
					base = atomicAdd (&s_out, (u32) __popc (bal));
				base = __shfl_sync (0xFFFFFFFFu, base, 0);
				if (keep)
				{
					const uint4 *src = reinterpret_cast<const uint4 *> (st + i * RECW);
					uint4 *dst = reinterpret_cast<uint4 *> (rec2 + (out + base + __popc (bal & ((1u << lane) - 1u))) * RECW);
#pragma unroll
					for (u32 q = 0; q < VEC; q++)
						dst[q] = src[q];
				}
			}
			__syncthreads ();
			out += s_out;
			__syncthreads ();	// s_out, the chunk and the table are rewritten by the next pass
		}
		if (tid == 0)
		{
			end[sl] = out;
			kept_total += out - r0;
		}
	}
	if (tid == 0 && kept_total)
		atomicAdd (n_kept, kept_total);
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

template <int W> struct SkmImage
{
	u64 *key;	// [S * W]
	u64 *ord;	// [S]
	u32 *cell;	// [CELL_WORDS * S]: word q of slot i at q * S + i
	u32 *extra;	// [S]
	u32 *state;	// [S] (W > 1): 0 empty, 1 key being written, 2 occupied
};

__host__ __device__ inline size_t skm_image_bytes (int W, u32 S)
{
	return (size_t) S * (8 * W + 8 + 4 * CELL_WORDS + 4 + (W > 1 ? 4 : 0));
}
__host__ __device__ inline size_t skm_build_smem (int W, const SkmGeom &g)
{	// image + window prefix of a chunk of records
	return skm_image_bytes (W, g.slice_slots) + 4 * (2 * (size_t) g.build_nt + 4);
}

// Double hashing (S is prime, 1 <= step < S): in shared memory a probe costs the same wherever it
// lands, and a warp step lasts as long as its longest probe sequence — at half load the longest of a
// few unsuccessful searches is ~4 probes here against ~13 with linear probing.
template <int W>
__device__ __forceinline__ u32 skm_find (const SkmImage<W> &im, u32 S, const Key<W> &key, u32 idx, u32 step)
{
	// a probe sequence that finds SKM_MAX_TRIES slots taken gives up: the image is as good as full
	// (0.9^128 = 1e-6), and the caller has the work item retried split by k-mer hash
	for (u32 tries = 0; tries < min (S, SKM_MAX_TRIES);)
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
					return idx;
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
		idx += step;
		if (idx >= S)
			idx -= S;
		tries++;
	}
	return S;
}

struct SkmWork { u32 slice, r, R, nrec; };	// keys of `slice` with sub-hash % R == r; nrec: records of the slice (set when an item fails)

static constexpr u32 MAX_SWEEPS = 4;	// slice_slots <= MAX_SWEEPS * threads per CTA

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

// A slice that overflowed its image is cut into q sub-slices by k-mer hash in ONE pass over its
// records: every window becomes a one-window sub-record (the window's K bases and its neighbours,
// the record's multiplicity, the window's ordinal) in the run of sub-slice qbase + bucket.  The
// sub-slices then go through skm_dedupe_kernel and skm_build_kernel like any slice.  (Retrying a
// slice as q work items that each filter the windows by hash scans its records q times; the slices
// that overflow are the highly expressed loci — most of the reads of a skewed data set.)
// PASS 0 counts sub-records per sub-slice (hist2), PASS 1 writes them (cur2 starts at the scanned offsets).
struct SkmSplit { u32 slice, q, qbase, rlo, rhi, pad[3]; };	// records [rlo, rhi) of the slice's run -> sub-slices qbase .. qbase + q
static constexpr int RS_NT = 256;

__device__ __forceinline__ u32 skm_bucket (u64 h, u32 q) { return __umulhi ((u32) (h >> 32) * 0x9E3779B1u, q); }

template <int W, int PASS>
__global__ void __launch_bounds__ (RS_NT)
skm_resplit_kernel (const u32 *rec2, const u64 *off, const u64 *end, const SkmSplit *splits, u32 n_splits, int K,
		    u32 *hist2, unsigned long long *cur2, u32 *rec3)
{
	constexpr u32 RECW = SkmRec<W>::WORDS, LAST = RECW - SKM_HDR - 1;
	Key<W> kmask;
#pragma unroll
	for (int q = 0; q < W; q++)
	{
		const int bits = 2 * K - 64 * (W - 1 - q);
		kmask.w[q] = bits >= 64 ? ~0ull : (bits > 0 ? (1ull << bits) - 1 : 0ull);
	}
	const int top = 2 * (K - 1);
	// one CTA per chunk of a failed slice's records (the host cuts big slices into chunks so that a
	// huge slice does not sit on one CTA)
	for (u32 sp = blockIdx.x; sp < n_splits; sp += gridDim.x)
	{
		const SkmSplit S = splits[sp];
		const u64 r0 = off[S.slice], r1 = min (end[S.slice], r0 + S.rhi);
		for (u64 i = r0 + S.rlo + threadIdx.x; i < r1; i += RS_NT)
		{
			const u32 *rec = rec2 + i * RECW;
			SkmRoll<W> st;
			skm_roll_init<W, false> (st, rec, K, 0);
			const u32 h1 = __ldg (rec + 1);
			const u32 phl = (h1 >> 14) & 1u, nrun = (h1 >> 15) & 1u, pnb = h1 >> 16;
			const u32 phr = nrun ? 0u : pnb - phl - (u32) K - (st.n - 1);
#pragma unroll 1
			for (u32 t = 0; t < st.n; t++)
			{
				Key<W> key;
				u32 left, right;
				skm_roll_window<W> (st, key, left, right);
				const u32 g = S.qbase + skm_bucket (key_hash<W> (key), S.q);
				if (PASS == 0)
					atomicAdd (hist2 + g, 1u);
				else
				{
					const u64 pos = atomicAdd (cur2 + g, 1ull);
					u32 *dst = rec3 + pos * RECW;
					const u32 hl = nrun ? 0u : (t > 0 || phl), hr = nrun ? 0u : (t + 1 < st.n || phr);
					const u32 nb = nrun ? 0u : hl + (u32) K + hr, first = nrun ? 0u : phl + t - hl;
					u32 wd[4];
					wd[0] = (u32) st.ord;
					wd[1] = (u32) (st.ord >> 32) | (hl << 14) | (nrun << 15) | (nb << 16);	// one window: n - 1 = 0
					wd[2] = st.add;	// multiplicity (an N-run: all of its windows)
					const u32 nbw = (nb + 15) >> 4;
					const u32 *rd = rec + SKM_HDR;
#pragma unroll
					for (u32 q = 0; q < RECW - SKM_HDR; q++)
					{
						u32 v = 0;
						if (q < nbw)
						{
							const u32 b = first + 16 * q, wq = b >> 4, sh = 2 * (b & 15);
							v = __funnelshift_l (__ldg (rd + min (wq + 1, LAST)), __ldg (rd + min (wq, LAST)), sh);
							if (q == nbw - 1 && (nb & 15))
								v &= 0xFFFFFFFFu << (32 - 2 * (nb & 15));
						}
						const u32 o = SKM_HDR + q;
						wd[o & 3] = v;
						if ((o & 3) == 3)
							*reinterpret_cast<uint4 *> (dst + (o & ~3u)) = make_uint4 (wd[0], wd[1], wd[2], wd[3]);
					}
				}
				skm_roll_step<W> (st, kmask, top);
			}
		}
	}
}

// One CTA per work item, items handed out through *item_cursor.  items == nullptr: item i is (slice i, 0, 1).
//
// Records -> image: the windows of up to 2 NT records are flattened and cut into NT equal runs,
// one per thread (a slice holds about one record per thread, of 1 to 32 windows: whole records per
// lane leave most lanes waiting for the longest).  A thread finds the record of its first window
// (binary search in the window prefix), sets up the rolling state there (one 16-byte header load +
// one extraction) and rolls on, re-seating itself when it crosses into the next record.  Instances
// that meet in the same (slot, cell) within a warp step are added by one lane for all.
// Image -> node store: every warp owns a contiguous range of slots; occupied slots are ranked by
// ballot + a scan of the 32 warp totals, thread 0 reserves the item's space in the store, and the
// nodes go out in slot order, 32 bytes per lane, consecutive lanes to consecutive nodes.
template <int W, int NT>
__global__ void __launch_bounds__ (NT, 1024 / NT)
skm_build_kernel (typename SlotOf<W>::type *store, u64 store_cap, unsigned long long *node_cursor, SkmGeom g, int K,
		  const u32 *rec2, const u64 *off, const u64 *end, const SkmWork *items, u32 n_items, unsigned long long *item_cursor,
		  SkmWork *failed, u32 *n_failed, u32 max_failed, Counters *ctr)
{
	typedef typename SlotOf<W>::type S_t;
	extern __shared__ __align__(16) u32 smem[];
	__shared__ u32 s_full, s_item[2], s_warp[NT / 32];
	__shared__ unsigned long long s_base;
	const u32 S = g.slice_slots, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const u32 spw = ((S + NT - 1) / NT) * 32;	// slots per warp in the compaction (a multiple of 32)
	SkmImage<W> im;
	im.key = reinterpret_cast<u64 *> (smem);
	im.ord = im.key + (size_t) S * W;
	im.cell = reinterpret_cast<u32 *> (im.ord + S);
	im.extra = im.cell + (size_t) S * CELL_WORDS;
	im.state = im.extra + S;
	u32 *pre = im.extra + S + (W > 1 ? S : 0);	// [2 NT + 1]: exclusive prefix of the windows of a chunk's records
	Key<W> kmask;	// the low 2K bits
#pragma unroll
	for (int q = 0; q < W; q++)
	{
		const int bits = 2 * K - 64 * (W - 1 - q);
		kmask.w[q] = bits >= 64 ? ~0ull : (bits > 0 ? (1ull << bits) - 1 : 0ull);
	}
	const int top = 2 * (K - 1);
	u64 nodes = 0, inst = 0;	// nodes: thread 0 only; inst: instances this thread applied (items that were written)
	if (tid == 0)
	{
		s_full = 0;
		s_item[0] = (u32) atomicAdd (item_cursor, 1ull);
	}
	for (u32 i = tid; i < S; i += NT)
	{
#pragma unroll
		for (int q = 0; q < W; q++)
			im.key[(size_t) i * W + q] = EMPTY64;
		im.ord[i] = ORD40_NONE;
		im.extra[i] = 0u;
		if constexpr (W > 1)
			im.state[i] = 0u;
	}
	for (u32 i = tid; i < S * CELL_WORDS; i += NT)
		im.cell[i] = 0u;
	__syncthreads ();
	for (u32 round = 0;; round++)
	{
		const u32 it = s_item[round & 1];
		if (it >= n_items)
			break;
		u32 next_item = 0;
		if (tid == 0)	// the next item is asked for now and looked at after this one: the round trip is hidden
			next_item = (u32) atomicAdd (item_cursor, 1ull);
		SkmWork wk;
		if (items)
			wk = items[it];
		else
		{
			wk.slice = it;
			wk.r = 0;
			wk.R = 1;
			wk.nrec = 0;
		}
		const u64 r0 = off[wk.slice], r1 = end[wk.slice];	// the run's surviving records (skm_dedupe_kernel)
		u64 mine = 0;
		for (u64 c0 = r0; c0 < r1 && !*reinterpret_cast<volatile u32 *> (&s_full); c0 += 2 * NT)
		{	// up to 2 NT records at a time: their windows are flattened (exclusive prefix in pre[]) and
			// cut into NT equal runs, one per thread
			const u32 nrec = (u32) min ((u64) (2 * NT), r1 - c0);
			const u32 *recs = rec2 + c0 * g.recw;
			u32 nw[2] = { 0, 0 };
#pragma unroll
			for (int q = 0; q < 2; q++)
				if (2 * tid + q < nrec)
				{
					const u32 h1 = __ldg (recs + (size_t) (2 * tid + q) * g.recw + 1);
					nw[q] = ((h1 >> 15) & 1u) ? 1u : ((h1 >> 8) & 63u) + 1;	// an N-run is applied in one go
				}
			u32 incl = nw[0] + nw[1];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1)
			{
				const u32 y = __shfl_up_sync (0xFFFFFFFFu, incl, d);
				if (lane >= (u32) d)
					incl += y;
			}
			if (lane == 31)
				s_warp[wid] = incl;
			__syncthreads ();
			u32 total;
			{
				const u32 c = lane < NT / 32 ? s_warp[lane] : 0u;
				u32 in2 = c;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1)
				{
					const u32 y = __shfl_up_sync (0xFFFFFFFFu, in2, d);
					if (lane >= (u32) d)
						in2 += y;
				}
				const u32 lower = __shfl_sync (0xFFFFFFFFu, in2 - c, wid);
				total = __shfl_sync (0xFFFFFFFFu, in2, 31);
				const u32 excl = lower + incl - nw[0] - nw[1];
				if (2 * tid < nrec)
					pre[2 * tid] = excl;
				if (2 * tid + 1 < nrec)
					pre[2 * tid + 1] = excl + nw[0];
				if (tid == 0)
					pre[nrec] = total;
			}
			__syncthreads ();
			const u32 per = (total + NT - 1) / NT;	// windows per thread
			const u32 w0 = tid * per, w1 = min (total, w0 + per);
			SkmRoll<W> st;
			st.n = st.t = 0;
			u32 x = 0;
			if (w0 < w1)
			{
				u32 lo = 0, hi = nrec - 1;	// record of window w0: largest x with pre[x] <= w0
				while (lo < hi)
				{
					const u32 mid = (lo + hi + 1) >> 1;
					if (pre[mid] <= w0)
						lo = mid;
					else
						hi = mid - 1;
				}
				x = lo;
				skm_roll_init<W> (st, recs + (size_t) x * g.recw, K, w0 - pre[x]);
			}
			for (u32 t = 0; t < per && !*reinterpret_cast<volatile u32 *> (&s_full); t++)
			{
				const bool act = w0 + t < w1;
				if (act && st.t == st.n)	// on to the next record
					skm_roll_init<W> (st, recs + (size_t) ++x * g.recw, K, 0);
				u32 idx = S, cellid = 0;
				bool wanted = false;
				if (act)
				{
					Key<W> key;
					u32 left, right;
					skm_roll_window<W> (st, key, left, right);
					const u64 h = key_hash<W> (key);
					wanted = wk.R == 1 || (u32) (h >> 32) % wk.R == wk.r;
					if (wanted)
						idx = skm_find<W> (im, S, key, home_of (h, S), 1u + __umulhi ((u32) (h >> 32), S - 1));
					cellid = left * 5 + right;
				}
				__syncwarp ();	// probe sequences differ in length: meet again before the update
				const bool hit = idx < S;
				// lanes of this step that meet in the same (slot, cell) with one instance each: the lowest one adds for all
				const bool one = hit && st.add == 1;
				const u32 peers = __match_any_sync (0xFFFFFFFFu, one ? idx * 32 + cellid : 0xFFFFFFFFu - lane);
				if (hit)
				{
					u32 *cw = im.cell + (cellid >> 1) * S + idx;
					const u32 sh = 16 * (cellid & 1);
					if (one)
					{
						if ((u32) (__ffs (peers) - 1) == lane)
						{
							const u32 cnt = __popc (peers);
							if (((*reinterpret_cast<volatile u32 *> (cw) >> sh) & 0xFFFFu) >= CELL_STOP)
								atomicAdd (im.extra + idx, cnt);
							else
								atomicAdd (cw, cnt << sh);
						}
					}
					else
					{	// a multiplicity: the cell takes what can still matter to a 6-bit link counter, `extra` the rest
						// (in flight at most 63 per thread of the CTA on top of 62: below 2^16)
						const u32 inc = ((*reinterpret_cast<volatile u32 *> (cw) >> sh) & 0xFFFFu) >= LINK_SAT ? 0u : min (st.add, LINK_SAT);
						if (inc)
							atomicAdd (cw, inc << sh);
						if (st.add > inc)
							atomicAdd (im.extra + idx, st.add - inc);
					}
					if (st.ord < *reinterpret_cast<volatile u64 *> (im.ord + idx))
						atomicMin (im.ord + idx, st.ord);
					mine += st.add;
				}
				else if (wanted)
					s_full = 1;
				if (act)
					skm_roll_step<W> (st, kmask, top);
			}
			__syncthreads ();	// pre[] and s_warp[] are rewritten by the next chunk
		}
		// ---- image -> node store
		const bool full = s_full != 0;
		// every warp lists the occupied slots of its range (pre[] is free now) and counts them
		unsigned short *wl = reinterpret_cast<unsigned short *> (pre) + wid * spw;
		u32 cnt = 0;
#pragma unroll
		for (u32 sw = 0; sw < MAX_SWEEPS; sw++)
			if (sw * 32 < spw)
			{
				const u32 i = wid * spw + sw * 32 + lane;
				bool occ = false;
				if (i < S)
				{
					if constexpr (W == 1)
						occ = im.key[i] != EMPTY64;
					else
						occ = im.state[i] == 2u;
				}
				const u32 bal = __ballot_sync (0xFFFFFFFFu, occ);
				if (occ)
					wl[cnt + __popc (bal & ((1u << lane) - 1u))] = (unsigned short) i;
				cnt += __popc (bal);
			}
		if (lane == 0)
			s_warp[wid] = cnt;
		__syncthreads ();
		u32 run, tot;
		{	// every warp scans the 32 warp totals
			const u32 c = lane < NT / 32 ? s_warp[lane] : 0u;
			u32 incl = c;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1)
			{
				const u32 y = __shfl_up_sync (0xFFFFFFFFu, incl, d);
				if (lane >= (u32) d)
					incl += y;
			}
			run = __shfl_sync (0xFFFFFFFFu, incl - c, wid);
			tot = __shfl_sync (0xFFFFFFFFu, incl, 31);
		}
		if (tid == 0)
		{
			unsigned long long b = 0;
			bool fail = full;
			if (!fail && tot)
			{
				b = atomicAdd (node_cursor, (unsigned long long) tot);
				if (b + tot > store_cap)
				{
					fail = true;
					atomicOr (reinterpret_cast<unsigned long long *> (&ctr->overflow), 4ull);	// node store exhausted
				}
			}
			if (full)
			{	// retried later, split by k-mer hash
				const u32 f = atomicAdd (n_failed, 1u);
				wk.nrec = (u32) min (r1 - r0, (u64) 0xFFFFFFFFu);
				if (f < max_failed)
					failed[f] = wk;
				else
					atomicOr (reinterpret_cast<unsigned long long *> (&ctr->overflow), 8ull);
			}
			s_base = fail ? ~0ull : b;
			if (!fail)
				nodes += tot;
		}
		__syncthreads ();
		const u64 nbase = s_base;
		const bool write = nbase != ~0ull;
		for (u32 k = lane; k < cnt; k += 32)
		{	// one lane per node: consecutive lanes write consecutive slots of the store
			const u32 i = wl[k];
			u32 row[4] = { 0, 0, 0, 0 }, col[4] = { 0, 0, 0, 0 }, count = im.extra[i];
#pragma unroll
			for (int q = 0; q < CELL_WORDS; q++)
			{
				const u32 v = im.cell[q * S + i];
				im.cell[q * S + i] = 0u;
#pragma unroll
				for (int hlf = 0; hlf < 2; hlf++)
				{
					const int c = 2 * q + hlf;
					if (c < 25)
					{
						const u32 x = hlf ? v >> 16 : v & 0xFFFFu;
						count += x;
						if (c / 5 < 4)
							row[c / 5] += x;	// five 16-bit terms: no overflow
						if (c % 5 < 4)
							col[c % 5] += x;
					}
				}
			}
			u32 L = 0, R = 0;
#pragma unroll
			for (int b = 0; b < 4; b++)
			{
				L |= min (row[b], LINK_SAT) << (6 * b);
				R |= min (col[b], LINK_SAT) << (6 * b);
			}
			const u64 w0 = (im.ord[i] << 24) | L, w1 = ((u64) count << 32) | R;
			Key<W> k2;
#pragma unroll
			for (int q = 0; q < W; q++)
			{
				k2.w[q] = im.key[(size_t) i * W + q];
				im.key[(size_t) i * W + q] = EMPTY64;
			}
			im.ord[i] = ORD40_NONE;
			im.extra[i] = 0u;
			if constexpr (W > 1)
				im.state[i] = 0u;
			if (write)
			{
				S_t *dst = store + nbase + run + k;
				if constexpr (W == 1)
					st256 (dst, k2.w[0], 0ull, w0, w1);
				else if constexpr (W == 2)
					st256 (dst, k2.w[0], k2.w[1], w0, w1);
				else
				{
					st256 (dst, k2.w[0], k2.w[1], k2.w[2], k2.w[3]);
					st256 (reinterpret_cast<u64 *> (dst) + 4, w0, w1, 0ull, 0ull);
				}
			}
		}
		if (write)	// instances applied by a work item that is going to be retried are not counted
			inst += mine;
		if (tid == 0)
		{
			s_full = 0;
			s_item[(round + 1) & 1] = next_item;
		}
		__syncthreads ();	// the image is clean; s_item[] of the next round is in place
	}
	{
#pragma unroll
		for (int d = 16; d > 0; d >>= 1)
			inst += __shfl_down_sync (0xFFFFFFFFu, inst, d);
		if (lane == 0 && inst)
			atomicAdd (&ctr->n_instances, inst);
	}
	if (tid == 0 && nodes)
		atomicAdd (&ctr->n_nodes, nodes);
}

}	// namespace sdt
