// sdt_build.cuh — skm_build_kernel: the slices of the table are built in shared memory, one work item
// (slice, or piece of a slice that overflowed) at a time per SM, and written to the node store.
//
// What the phases of a slice cost (measured with phase clocks on C2, round 2): the upserts are bound
// by the shared-memory atomic pipe (2 cycles per lane, a 64-bit CAS twice that: B300_MICROARCH.md), the
// compaction by shared-memory traffic, and ~a quarter of a slice's time went to the chain of
// dependent global round trips in front of them (work-item cursor -> run offsets -> record headers
// -> records).  Splitting the CTA into two groups that take turns on the image (fetch of one slice
// under the upserts of the other) was measured and is SLOWER (49.8 vs 42.6 ms): the upserts are
// latency-bound per warp and want all 32 warps.  So all 1024 threads work on one slice, and the round
// trips are taken off the critical path instead:
//   * work items are drawn three rounds ahead, their run offsets are loaded two rounds ahead (one
//     thread each), and the records of the NEXT item are copied into a second staging area with
//     cp.async while this item is processed; headers, window prefix, seats and re-seats read shared memory;
//   * the node-store reservation (one global atomic per slice) is issued as soon as the slice's
//     node count is known and travels while the cells are summed; the finished payloads wait in the
//     image (cell words 0-3) for it;
//   * ordinals are kept in 32 bits when everything pushed fits (all BASELINE configs do): a native
//     ATOMS.MIN instead of a 64-bit CAS loop, and 4 bytes less per slot;
//   * home slot and probe step come from a 32-bit hash (two multiplies per key word instead of fmix64).
#pragma once
#include "sdt_merge.cuh"

namespace sdt {

static constexpr int BUILD_NT = 1024, BUILD_NW = BUILD_NT / 32;
template <int W> struct BuildCfg { static constexpr u32 RCAP = W == 1 ? 512u : (W == 2 ? 384u : 256u); };	// records per staged chunk (<= BUILD_NT)

__host__ __device__ inline u32 skm_rcap (int W) { return W == 1 ? 512u : (W == 2 ? 384u : 256u); }
__host__ __device__ inline u32 skm_recw (int W) { return W == 1 ? 8u : (W == 2 ? 12u : 16u); }
__host__ __device__ inline size_t skm_image2_bytes (int W, u32 S, bool ord32)
{	// key, ordinal, 5 x 5 cells, overflow count, state (multi-word keys)
	return (((size_t) S * (8 * W + (ord32 ? 4 : 8) + 4 * CELL_WORDS + 4 + (W > 1 ? 4 : 0))) + 15) & ~(size_t) 15;
}
__host__ __device__ inline size_t skm_stage_words (int W) { return (size_t) skm_rcap (W) * skm_recw (W); }
static constexpr u32 BUILD_WL_BYTES = 8192;	// the compaction's lists of occupied slots: 32 warps x <= 128 slots x 2 bytes
__host__ __device__ inline size_t skm_build2_smem (int W, u32 S, bool ord32)
{	// image + staging area + slot lists + window prefix
	return skm_image2_bytes (W, S, ord32) + 4 * (skm_stage_words (W) + skm_rcap (W) + 4) + BUILD_WL_BYTES;
}
// largest image that fits beside them (227 KB per CTA, ~1 KB static); sized for 64-bit ordinals so that
// the slice geometry does not depend on how many reads are going to be pushed
inline u32 skm_build2_max_slots (int W)
{
	const size_t avail = 227 * 1024 - 1024 - 4 * (skm_stage_words (W) + skm_rcap (W) + 4) - BUILD_WL_BYTES - 16;
	return (u32) std::min<size_t> (4096, avail / (8 * W + 8 + 4 * CELL_WORDS + 4 + (W > 1 ? 4 : 0)));
}

#ifdef SDT_BUILD_PROF
__device__ unsigned long long g_build_prof[8];	// phase clocks (thread 0): prepare, -, insert, compact, items, chunks
#define PROF_MARK(q) do { const long long t1_ = clock64 (); prof[q] += t1_ - tp; tp = t1_; } while (0)
#else
#define PROF_MARK(q) do { } while (0)
#endif

__device__ __forceinline__ void cp_async16 (void *smem_dst, const void *gsrc)
{
	const u32 d = (u32) __cvta_generic_to_shared (smem_dst);
	asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit () { asm volatile ("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait () { asm volatile ("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// rolling state at window tw of a record that sits in shared memory
template <int W>
__device__ __forceinline__ void skm_seat (SkmRoll<W> &s, const u32 *rec, int K, u32 tw)
{
	constexpr u32 LAST = SkmRec<W>::WORDS - SKM_HDR - 1;
	const uint4 hd = *reinterpret_cast<const uint4 *> (rec);	// ord low | header | multiplicity | first bases
	const u32 h1 = hd.y;
	const u32 n = ((h1 >> 8) & 63u) + 1, has_left = (h1 >> 14) & 1u, nrun = (h1 >> 15) & 1u, nb = h1 >> 16;
	s.ord = ((u64) hd.x | ((u64) (h1 & 0xFFu) << 32)) + tw;
	s.t = tw;
	if (nrun)
	{	// n instances of key 0 without links: one step
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
	s.n = n;
	s.has_right = nb - has_left - (u32) K - (n - 1);
	s.add = hd.z;
	if constexpr (W == 1)
		if (tw == 0)
		{	// start of a record (every re-seat): 160 bits of bases in two 16-byte loads
			const uint4 h2 = *(reinterpret_cast<const uint4 *> (rec) + 1);
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
		s.pend[1] = n - tw > 32 ? bases64 (rd, p0 + 32, LAST) : 0ull;
}

template <int W, bool ORD32> struct SkmImage2
{
	typedef typename std::conditional<ORD32, u32, u64>::type ord_t;
	u64 *key;	// [S * W]
	ord_t *ord;	// [S]
	u32 *cell;	// [CELL_WORDS * S]: word q of slot i at q * S + i
	u32 *extra;	// [S]
	u32 *state;	// [S] (W > 1): 0 empty, 1 key being written, 2 occupied
};

// probe of the image by double hashing (S is prime, 1 <= step < S): a probe costs the same wherever it
// lands, and a warp step lasts as long as its longest probe sequence.  A sequence that finds
// SKM_MAX_TRIES slots taken gives up: the image is as good as full and the item is retried split by hash.
template <int W, bool ORD32>
__device__ __forceinline__ u32 skm_find2 (const SkmImage2<W, ORD32> &im, u32 S, const Key<W> &key, u32 idx, u32 step)
{
	const u32 lim = min (S, SKM_MAX_TRIES);
	for (u32 tries = 0; tries < lim;)
	{
		if constexpr (W == 1)
		{
			u64 k = *reinterpret_cast<volatile u64 *> (im.key + idx);
			if (k == key.w[0])
				return idx;
			if (k == EMPTY64)
			{
				k = atomicCAS (im.key + idx, EMPTY64, key.w[0]);
				if (k == EMPTY64 || k == key.w[0])
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

struct SkmDesc { u64 r0, r1; u32 it, wsum, r, R; };	// a work item with its run of records

// One CTA per SM, work items handed out through *item_cursor.
template <int W, bool ORD32>
__global__ void __launch_bounds__ (BUILD_NT, 1)
skm_build2_kernel (typename SlotOf<W>::type *store, u64 store_cap, unsigned long long *node_cursor, SkmGeom g, int K,
		   const u32 *rec2, const SkmWork *items, u32 n_items, const unsigned long long *n_items_dev, unsigned long long *item_cursor,
		   SkmWork *failed, u32 *n_failed, u32 max_failed, Counters *ctr)
{
	typedef typename SlotOf<W>::type S_t;
	typedef typename SkmImage2<W, ORD32>::ord_t ord_t;
	constexpr u32 RECW = SkmRec<W>::WORDS, RCAP = BuildCfg<W>::RCAP, VEC = RECW / 4;
	constexpr ord_t ORD_NONE = ORD32 ? (ord_t) 0xFFFFFFFFu : (ord_t) ORD40_NONE;
	extern __shared__ __align__(16) u32 smem[];
	__shared__ u32 s_full, s_it[4], s_warp[BUILD_NW];
	__shared__ SkmDesc s_desc[4];
	__shared__ unsigned long long s_base;
	const u32 S = g.slice_slots, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const u32 spw = ((S + BUILD_NT - 1) / BUILD_NT) * 32;	// slots per warp in the compaction (a multiple of 32)
	if (n_items_dev)	// the number of work items as skm_merge_kernel left it on the device (the host has not looked)
		n_items = (u32) min ((unsigned long long) n_items, *n_items_dev);
	SkmImage2<W, ORD32> im;
	im.key = reinterpret_cast<u64 *> (smem);
	im.ord = reinterpret_cast<ord_t *> (im.key + (size_t) S * W);
	im.cell = reinterpret_cast<u32 *> (im.ord + S);
	im.extra = im.cell + (size_t) S * CELL_WORDS;
	im.state = im.extra + S;
	u32 *stg = smem + skm_image2_bytes (W, S, ORD32) / 4;	// staging area of RCAP records
	unsigned short *wl0 = reinterpret_cast<unsigned short *> (stg + RCAP * RECW);	// slot lists of the compaction
	u32 *pre = stg + RCAP * RECW + BUILD_WL_BYTES / 4;	// [RCAP + 1] exclusive prefix of the windows of a chunk's records
	Key<W> kmask;	// the low 2K bits
#pragma unroll
	for (int q = 0; q < W; q++)
	{
		const int bits = 2 * K - 64 * (W - 1 - q);
		kmask.w[q] = bits >= 64 ? ~0ull : (bits > 0 ? (1ull << bits) - 1 : 0ull);
	}
	const int top = 2 * (K - 1);
#ifdef SDT_BUILD_PROF
	long long prof[6] = { 0, 0, 0, 0, 0, 0 }, tp = clock64 ();
#endif
	u64 nodes = 0, inst = 0;	// nodes: thread 0 only; inst: instances this thread applied (items that were written)

	// a work item's descriptor (one thread, one load)
	auto load_desc = [&](u32 it, SkmDesc &d) {
		d.it = it;
		d.wsum = 0; d.r = 0; d.R = 1;
		d.r0 = d.r1 = 0;
		if (it < n_items)
		{
			const SkmWork wk = items[it];
			d.wsum = wk.wsum; d.r = wk.r; d.R = wk.R;
			d.r0 = wk.r0;
			d.r1 = wk.r0 + wk.nrec;
		}
	};
	// the first chunk of an item's records into a staging area (asynchronously; one 16-byte piece per thread and turn)
	auto prefetch = [&](const SkmDesc &d) {
		const u32 nrec = (u32) min ((u64) RCAP, d.r1 - d.r0);
		const uint4 *src = reinterpret_cast<const uint4 *> (rec2 + d.r0 * RECW);
		uint4 *dst = reinterpret_cast<uint4 *> (stg);
		for (u32 v = tid; v < nrec * VEC; v += BUILD_NT)
			cp_async16 (dst + v, src + v);
		cp_async_commit ();
	};

	if (tid == 0)
	{
		s_full = 0;
		for (int q = 0; q < 3; q++)
			s_it[q] = (u32) atomicAdd (item_cursor, 1ull);
		SkmDesc d;
		load_desc (s_it[0], d);
		s_desc[0] = d;
		load_desc (s_it[1], d);
		s_desc[1] = d;
	}
	for (u32 i = tid; i < S; i += BUILD_NT)
	{
#pragma unroll
		for (int q = 0; q < W; q++)
			im.key[(size_t) i * W + q] = EMPTY64;
		im.ord[i] = ORD_NONE;
		im.extra[i] = 0u;
		if constexpr (W > 1)
			im.state[i] = 0u;
	}
	for (u32 i = tid; i < S * CELL_WORDS; i += BUILD_NT)
		im.cell[i] = 0u;
	__syncthreads ();
	prefetch (s_desc[0]);
	for (u32 round = 0;; round++)
	{
		const SkmDesc D = s_desc[round & 3];
		if (D.it >= n_items)
			break;
		// ---- the pipeline in front of this item: cursor three rounds ahead, descriptor two (both issued in front of
		// the compaction, further down), records one: they were copied under the previous item's compaction
		cp_async_wait<0> ();
		__syncthreads ();
		u64 mine = 0;
		if (D.R == 0 && tid == 0)
			s_full = 1;	// a chain far beyond an image: straight to the failed list (it is cut into sub-slices)
		if (D.R == 0)
			__syncthreads ();
		for (u64 c0 = D.r0; c0 < D.r1 && !*reinterpret_cast<volatile u32 *> (&s_full); c0 += RCAP)
		{	// up to RCAP records at a time: their windows are flattened (exclusive prefix in pre[]) and
			// cut into 1024 equal runs, one per thread
			const u32 nrec = (u32) min ((u64) RCAP, D.r1 - c0);
			if (c0 != D.r0)
			{	// a slice with more records than a staging area holds: the further chunks are fetched in line
				const uint4 *src = reinterpret_cast<const uint4 *> (rec2 + c0 * RECW);
				uint4 *dst = reinterpret_cast<uint4 *> (stg);
				for (u32 v = tid; v < nrec * VEC; v += BUILD_NT)
					dst[v] = ldg_stream (src + v);
				__syncthreads ();
			}
			u32 nw = 0;
			if (tid < nrec)
			{
				const u32 h1 = stg[tid * RECW + 1];
				nw = ((h1 >> 15) & 1u) ? 1u : ((h1 >> 8) & 63u) + 1;	// an N-run is applied in one go
			}
			u32 incl = nw;
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
				const u32 c = s_warp[lane];
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
				if (tid < nrec)
					pre[tid] = lower + incl - nw;
				if (tid == 0)
					pre[nrec] = total;
			}
			__syncthreads ();
			const u32 per = (total + BUILD_NT - 1) / BUILD_NT;	// windows per thread
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
				skm_seat<W> (st, stg + (size_t) x * RECW, K, w0 - pre[x]);
			}
			PROF_MARK (0);
			for (u32 t = 0; t < per; t++)
			{
				if (__any_sync (0xFFFFFFFFu, *reinterpret_cast<volatile u32 *> (&s_full) != 0u))
					break;	// (warp-uniform: the loop body holds full-mask warp primitives)
				const bool act = w0 + t < w1;
				if (act && st.t == st.n)	// on to the next record
					skm_seat<W> (st, stg + (size_t) ++x * RECW, K, 0);
				u32 idx = S, cellid = 0;
				bool wanted = false;
				if (act)
				{
					Key<W> key;
					u32 left, right, ha, hb;
					skm_roll_window<W> (st, key, left, right);
					skm_hash2<W> (key, ha, hb);
					wanted = D.R == 1 || hb % D.R == D.r;
					if (wanted)
						idx = skm_find2<W, ORD32> (im, S, key, __umulhi (ha, S), 1u + __umulhi (hb, S - 1));
					cellid = left * 5 + right;
				}
				__syncwarp ();	// probe sequences differ in length: meet again before the update
				const bool hit = idx < S;
				// (lanes of a step that meet in the same (slot, cell) are left to the atomic unit: after the merge pass
				// records are distinct, such meetings are rare, and finding them with match.any cost every step 10 %)
				if (hit)
				{
					u32 *cw = im.cell + (cellid >> 1) * S + idx;
					const u32 sh = 16 * (cellid & 1);
					if (st.add == 1)
					{
						if (((*reinterpret_cast<volatile u32 *> (cw) >> sh) & 0xFFFFu) >= CELL_STOP)
							atomicAdd (im.extra + idx, 1u);
						else
							atomicAdd (cw, 1u << sh);
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
					const ord_t o = (ord_t) st.ord;
					if (o < *reinterpret_cast<volatile ord_t *> (im.ord + idx))
						atomicMin (im.ord + idx, o);
					mine += st.add;
				}
				else if (wanted)
					s_full = 1;
				if (act)
					skm_roll_step<W> (st, kmask, top);
			}
			__syncthreads ();	// pre[], s_warp[] and the staging area are rewritten by the next chunk; s_full is settled
			PROF_MARK (2);
#ifdef SDT_BUILD_PROF
			prof[5]++;
#endif
		}
		// ---- the next item's records travel while this one's image is compacted
		prefetch (s_desc[(round + 1) & 3]);
		// ---- image -> node store
		const bool full = s_full != 0;
		// (the cursor and the descriptor loads are issued here, where few registers are live, and land under the compaction)
		u32 it3 = 0;
		SkmDesc d2;
		if (tid == 0)
			it3 = (u32) atomicAdd (item_cursor, 1ull);
		if (tid == 32)
			load_desc (s_it[(round + 2) & 3], d2);
		// every warp lists the occupied slots of its range and counts them
		unsigned short *wl = wl0 + wid * spw;
		u32 cnt = 0;
		if (D.r1 > D.r0 && D.R != 0)
			for (u32 sw = 0; sw < spw; sw += 32)
			{
				const u32 i = wid * spw + sw + lane;
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
			const u32 c = s_warp[lane];
			u32 in2 = c;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1)
			{
				const u32 y = __shfl_up_sync (0xFFFFFFFFu, in2, d);
				if (lane >= (u32) d)
					in2 += y;
			}
			run = __shfl_sync (0xFFFFFFFFu, in2 - c, wid);
			tot = __shfl_sync (0xFFFFFFFFu, in2, 31);
		}
		unsigned long long resv = 0;
		if (tid == 0 && !full && tot)	// the reservation travels while the payloads are summed up
			resv = atomicAdd (node_cursor, (unsigned long long) tot);
		for (u32 k = lane; k < cnt; k += 32)
		{	// payload of every node, parked in cell words 0-3 of its slot
			const u32 i = wl[k];
			u32 row[4] = { 0, 0, 0, 0 }, col[4] = { 0, 0, 0, 0 }, count = im.extra[i];
			u32 cw[CELL_WORDS];
#pragma unroll
			for (int q = 0; q < CELL_WORDS; q++)
				cw[q] = im.cell[q * S + i];
#pragma unroll
			for (int q = 0; q < CELL_WORDS; q++)
			{
#pragma unroll
				for (int hlf = 0; hlf < 2; hlf++)
				{
					const int c = 2 * q + hlf;
					if (c < 25)
					{
						const u32 xx = hlf ? cw[q] >> 16 : cw[q] & 0xFFFFu;
						count += xx;
						if (c / 5 < 4)
							row[c / 5] += xx;	// five 16-bit terms: no overflow
						if (c % 5 < 4)
							col[c % 5] += xx;
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
			const u64 o40 = ORD32 ? (u64) im.ord[i] : (u64) im.ord[i];
			const u64 w0 = (o40 << 24) | L;
			im.cell[i] = (u32) w0;
			im.cell[S + i] = (u32) (w0 >> 32);
			im.cell[2 * S + i] = R;
			im.cell[3 * S + i] = count;
#pragma unroll
			for (int q = 4; q < CELL_WORDS; q++)
				im.cell[q * S + i] = 0u;
			im.extra[i] = 0u;
			im.ord[i] = ORD_NONE;
		}
		if (tid == 0)
		{
			bool fail = full;
			if (!fail && tot && resv + tot > store_cap)
			{
				fail = true;
				atomicOr (reinterpret_cast<unsigned long long *> (&ctr->overflow), 4ull);	// node store exhausted
			}
			if (full)
			{	// retried later, split by k-mer hash
				const u32 f = atomicAdd (n_failed, 1u);
				SkmWork wk;
				wk.r0 = D.r0; wk.wsum = D.wsum; wk.r = D.r; wk.R = D.R;
				wk.nrec = (u32) (D.r1 - D.r0);
				if (f < max_failed)
					failed[f] = wk;
				else
					atomicOr (reinterpret_cast<unsigned long long *> (&ctr->overflow), 8ull);
			}
			s_base = fail ? ~0ull : resv;
			if (!fail)
				nodes += tot;
			s_full = 0;
			s_it[(round + 3) & 3] = it3;
		}
		if (tid == 32)
			s_desc[(round + 2) & 3] = d2;
		__syncthreads ();
		const u64 nbase = s_base;
		const bool write = nbase != ~0ull;
		for (u32 k = lane; k < cnt; k += 32)
		{	// one lane per node: consecutive lanes write consecutive slots of the store
			const u32 i = wl[k];
			const u64 w0 = ((u64) im.cell[S + i] << 32) | im.cell[i], w1 = ((u64) im.cell[3 * S + i] << 32) | im.cell[2 * S + i];
			Key<W> k2;
#pragma unroll
			for (int q = 0; q < W; q++)
			{
				k2.w[q] = im.key[(size_t) i * W + q];
				im.key[(size_t) i * W + q] = EMPTY64;
			}
#pragma unroll
			for (int q = 0; q < 4; q++)
				im.cell[q * S + i] = 0u;
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
		__syncthreads ();	// the image is clean
		PROF_MARK (3);
#ifdef SDT_BUILD_PROF
		prof[4]++;
#endif
	}
	cp_async_wait<0> ();
	{
#pragma unroll
		for (int d = 16; d > 0; d >>= 1)
			inst += __shfl_down_sync (0xFFFFFFFFu, inst, d);
		if (lane == 0 && inst)
			atomicAdd (&ctr->n_instances, inst);
	}
	if (tid == 0 && nodes)
		atomicAdd (&ctr->n_nodes, nodes);
#ifdef SDT_BUILD_PROF
	if (tid == 0)
		for (int q = 0; q < 6; q++)
			atomicAdd (&g_build_prof[q], (unsigned long long) prof[q]);
#endif
}

}	// namespace sdt
