// sdt_merge.cuh — what happens to the chains of records (sdt_chain.cuh) after the reads are through:
// skm_merge_kernel (copies collapse, every chain becomes a contiguous run), skm_append_kernel (records
// that arrive from other GPUs -> chains), skm_split_kernel (a slice that overflowed by far -> sub-slices).
#pragma once
#include "sdt_skm.cuh"

namespace sdt {

// ------------------------------------------------------------------------------------------------
// Copies of the same super-k-mer collapse into one record, and chains become the work items of the build.
// At the coverage of a transcriptome most records of a slice are byte-identical copies (every
// error-free read that spans a super-k-mer emits the same bases, neighbours and window count; only
// the ordinal differs), and a window costs the build ~20x what finding a copy costs here.  A
// surviving record carries its multiplicity in word 2 and the smallest ordinal of its copies: window
// t of every copy is the same (key, left, right) instance with ordinal ord0 + t, so count and link
// counters take the multiplicity (update_kmer, newhash.c:71-96, is a sum) and the node's first
// ordinal the minimum.
//
// A CTA walks a contiguous range of chains in GROUPS: as many consecutive chains as fit one chunk (found
// by warp 0 from the chains' record counts, one group ahead, while the current group is merged).  The
// group's records are staged together (16 bytes at a time, swizzled against bank conflicts, mg_swz), a
// table of record indices keyed by the record's content finds the copies, and the survivors are written
// out IN ORDER (block scan), so that consecutive chains are adjacent.  Then the CTA cuts the stream of
// chains into WORK ITEMS for skm_build_kernel at chain boundaries — an open item is carried from group
// to group: as many consecutive chains as have at most `budget` windows left between them.  Every
// instance of a k-mer is in one chain, so an item holds all instances of its k-mers; an item has at most
// as many distinct k-mers as windows, so with budget < image slots it CANNOT overflow an image, whatever
// the number of slices was guessed to be (too many slices only make the chains short).  Only a single
// chain with more windows than that — a highly expressed locus — becomes an item that the build may
// have to split (R = 0 tells it not even to try when the chain is far beyond an image).
// The CTA's space in out[] is reserved once, by its chains' record count (an upper bound), before the
// records are read.  Several GPUs, sending side (per_owner): the chains of an owner in pieces of MG_PIECE,
// every group's survivors into the owner's region, no work items.
static constexpr int MG_NT = 512;	// 2 CTAs of 80-96 KB per SM
static constexpr u32 MG_GMAX = 32;	// chains per group at most
static constexpr u32 MG_PIECE = 128;	// several GPUs, sending side: consecutive chains of one owner that a CTA takes at a time
static constexpr u32 MG_ITEMS = 16;	// work items a CTA collects before it asks for their places in the list
#ifndef SDT_MERGE_CHUNK1
#define SDT_MERGE_CHUNK1 2048u
#endif
template <int W> struct MergeCfg { static constexpr u32 CHUNK = W == 1 ? SDT_MERGE_CHUNK1 : 1024u, TABLE = 2 * CHUNK; };
template <int W> __host__ __device__ inline size_t skm_merge_smem () { return (size_t) MergeCfg<W>::CHUNK * SkmRec<W>::WORDS * 4 + (size_t) MergeCfg<W>::TABLE * 4 + 2 * (MergeCfg<W>::CHUNK / CH_BLK + MG_GMAX + 4) * 4 + MergeCfg<W>::CHUNK; }
// Where the 16-byte part `part` of staged record `rec` sits in the chunk.  Threads of a warp work on consecutive
// records; with records of 32 (64) bytes laid out plainly every access would meet 2 (4) of its quarter-warp in the
// same banks (and word-wise accesses 8 to 32: the shared-memory pipe was the kernel's busiest unit), so the parts
// of every other group of records are swapped around.  Records are only ever touched 16 bytes at a time.
template <int W> __device__ __forceinline__ u32 mg_swz (u32 rec, u32 part)
{
	constexpr u32 VEC = SkmRec<W>::WORDS / 4;
	if (W == 1)
		return rec * VEC + (part ^ ((rec >> 2) & 1u));
	if (W == 4)
		return rec * VEC + (part ^ ((rec >> 1) & 3u));
	return rec * VEC + part;	// 48 bytes: no two of eight consecutive records share a bank
}

struct MergeOut
{
	u32 *out;			// merged records
	unsigned long long *out_cursor;	// next free record of out[]
	u64 out_cap;
	SkmWork *items;			// work items of the build
	unsigned long long *n_items;
	u64 max_items;
	unsigned long long *n_kept;	// [0] records left, [1] their windows
	u32 budget, oversize;		// windows per item at most; a single chain with more windows than `oversize` is marked R = 0
	// multi-GPU send side: chains are global slices, `per_owner` of them per rank; the survivors of owner r are
	// packed without gaps from region[r] on (cursor rcur[r]) and carry their slice in the last word; no items
	u32 per_owner, tagged;	// tagged: the records carry their slice already (a chain holds several slices)
	const u64 *region;
	unsigned long long *rcur;
};

// HAS_MULT: word 2 of the incoming records already is a multiplicity (sub-records, records merged before an exchange)
template <int W, bool HAS_MULT>
__global__ void __launch_bounds__ (MG_NT, 2)
skm_merge_kernel (SkmChains ch, const u64 *boff, const u32 *blist, MergeOut mo)
{
	constexpr u32 RECW = SkmRec<W>::WORDS, CH = MergeCfg<W>::CHUNK, TS = MergeCfg<W>::TABLE, VEC = RECW / 4, RPT = CH / MG_NT;
	extern __shared__ __align__(16) u32 smem[];
	__shared__ u32 s_scan[MG_NT / 32], s_tot, s_wtot;
	__shared__ u32 s_nraw2[2 * MG_GMAX], s_rs[MG_GMAX + 1], s_pos[MG_GMAX + 1], s_wpre[MG_GMAX + 1];	// per chain of the group: records (this group's and the next one's), first record in the chunk, first survivor, windows before it
	__shared__ u64 s_b02[2 * MG_GMAX];
	__shared__ u32 s_bs[MG_GMAX + 1];	// blocks in front of every chain of the group
	__shared__ u32 s_traw[2], s_nch[2];	// records and chains of this group and of the next one
	__shared__ SkmWork s_items[MG_ITEMS];
	u32 *s_nraw = s_nraw2;	// (the first group's)
	u64 *s_b0 = s_b02;
	__shared__ unsigned long long s_start, s_ibase;
	u32 *st = smem;			// [CH * RECW]: the chunk's records
	u32 *tab = smem + CH * RECW;	// [TS]: record index + 1, 0 = free
	u32 *sbl = tab + TS;		// [CH / CH_BLK + G]: the chunk's blocks
	u32 *sbd = sbl + CH / CH_BLK + MG_GMAX + 4;	// [same]: where a block's records go in st[] | how many << 16
	unsigned char *flg = reinterpret_cast<unsigned char *> (sbd + CH / CH_BLK + MG_GMAX + 4);	// [CH]: after the dedupe, 0 = the record is a copy and goes, else its windows
	uint4 *st4 = reinterpret_cast<uint4 *> (st);
	const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const u32 owners = mo.per_owner ? (ch.n_chains + mo.per_owner - 1) / mo.per_owner : 1;
	const u32 span = mo.per_owner ? mo.per_owner : ch.n_chains;	// groups do not straddle owners
	const u32 ppo = (span + MG_PIECE - 1) / MG_PIECE;	// pieces per owner
	const u32 n_pieces = ppo * owners;	// (32-bit on purpose: a 64-bit division per group and thread was a fifth of the kernel's instructions)
	u64 kept_total = 0, win_total = 0;	// thread 0

	// dedupe of the nrec records staged in st[]
	auto dedupe = [&](u32 nrec) {
		for (u32 i = tid; i < nrec; i += MG_NT)
		{
			uint4 r[VEC];
#pragma unroll
			for (u32 q = 0; q < VEC; q++)
				r[q] = st4[mg_swz<W> (i, q)];
			const u32 h1 = r[0].y;
			if ((h1 >> 15) & 1u)
			{	// an N-run stays as it is
				flg[i] = 1;
				continue;
			}
			const u32 nw = ((h1 >> 8) & 63u) + 1;
			u32 hsh = h1 & ~0xFFu;
			hsh = (hsh ^ r[0].w) * 0x9E3779B1u + (hsh >> 15);
#pragma unroll
			for (u32 q = 1; q < VEC; q++)
			{
				hsh = (hsh ^ r[q].x) * 0x9E3779B1u + (hsh >> 15);
				hsh = (hsh ^ r[q].y) * 0x9E3779B1u + (hsh >> 15);
				hsh = (hsh ^ r[q].z) * 0x9E3779B1u + (hsh >> 15);
				hsh = (hsh ^ r[q].w) * 0x9E3779B1u + (hsh >> 15);
			}
			u32 slot = fmix32 (hsh) & (TS - 1);
			for (;;)
			{
				u32 e = tab[slot];
				if (e == 0u)
					e = atomicCAS (tab + slot, 0u, i + 1);
				if (e == 0u)
				{	// first of its kind
					flg[i] = (unsigned char) nw;
					break;
				}
				uint4 *rep = st4 + mg_swz<W> (e - 1, 0);
				const uint4 c = *rep;	// (words 0-2 of a representative change under it: only the constant bits are compared)
				bool same = ((c.y ^ h1) & ~0xFFu) == 0u && c.w == r[0].w;	// the low 8 bits are ordinal bits and change
#pragma unroll
				for (u32 q = 1; q < VEC; q++)
				{
					const uint4 d = st4[mg_swz<W> (e - 1, q)];
					same &= d.x == r[q].x && d.y == r[q].y && d.z == r[q].z && d.w == r[q].w;
				}
				if (same)
				{	// words 0-1 as one 64-bit number: the header bits above the ordinal are equal, so the minimum is the ordinal's
					atomicAdd (reinterpret_cast<u32 *> (rep) + 2, HAS_MULT ? r[0].z : 1u);
					const u64 mine = (u64) r[0].x | ((u64) r[0].y << 32);
					if (mine < *reinterpret_cast<volatile u64 *> (rep))
						atomicMin (reinterpret_cast<unsigned long long *> (rep), mine);
					flg[i] = 0;	// dropped
					break;
				}
				slot = (slot + 1) & (TS - 1);
			}
		}
	};
	// survivors of st[0 .. nrec) to out[base ..] in order; returns their number, s_wtot their windows;
	// the survivors / windows in front of the records listed in s_rs[0 .. nb] go to s_pos / s_wpre
	auto write_out = [&](u32 nrec, u64 base, u32 nb, u32 slice0) -> u32 {
		u32 packed = 0;	// thread t: records RPT * t .. RPT * t + RPT - 1; windows << 12 | records
#pragma unroll
		for (u32 k = 0; k < RPT; k++)
		{
			const u32 i = RPT * tid + k;
			const u32 f = i < nrec ? flg[i] : 0u;
			packed += (f << 12) | (f != 0u);
		}
		u32 incl = packed;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const u32 y = __shfl_up_sync (0xFFFFFFFFu, incl, d);
			if (lane >= (u32) d)
				incl += y;
		}
		if (lane == 31)
			s_scan[wid] = incl;
		__syncthreads ();
		u32 lower = 0;
		{
			const u32 c = lane < MG_NT / 32 ? s_scan[lane] : 0u;
			u32 in2 = c;
#pragma unroll
			for (int d = 1; d < MG_NT / 32; d <<= 1)
			{
				const u32 y = __shfl_up_sync (0xFFFFFFFFu, in2, d);
				if (lane >= (u32) d)
					in2 += y;
			}
			lower = __shfl_sync (0xFFFFFFFFu, in2 - c, wid);
			const u32 total = __shfl_sync (0xFFFFFFFFu, in2, MG_NT / 32 - 1);
			if (tid == 0)
			{
				s_tot = total & 0xFFFu;
				s_wtot = total >> 12;
			}
		}
		tab[tid] = lower + incl - packed;	// survivors / windows in front of this thread's first record (the table is free after the dedupe)
		__syncthreads ();
		// the copy runs over the records in staging order, a warp on 32 consecutive ones (no bank conflicts, and
		// consecutive survivors go to consecutive places): a record's place is its scan thread's plus the survivors
		// among that thread's earlier records
#pragma unroll
		for (u32 k = 0; k < RPT; k++)
		{
			const u32 i = k * MG_NT + tid;
			if (i < nrec && flg[i])
			{
				const u32 t = i / RPT;
				u32 below = 0;
				for (u32 k2 = RPT * t; k2 < i; k2++)
					below += flg[k2] != 0;
				const u64 pos = base + (tab[t] & 0xFFFu) + below;
				if (pos < mo.out_cap)
				{
					uint4 *dst = reinterpret_cast<uint4 *> (mo.out + pos * RECW);
#pragma unroll
					for (u32 q = 0; q < VEC; q++)
					{
						uint4 x = st4[mg_swz<W> (i, q)];
						if (mo.per_owner && !mo.tagged && q == VEC - 1)
						{	// the last base word is never used (80 / 144 / 208 bases of room for 64 / 128 / 192): the slice travels there
							u32 g = 0;
							while (g + 1 < nb && s_rs[g + 1] <= i)
								g++;
							x.w = slice0 + g;
						}
						dst[q] = x;
					}
				}
			}
		}
		if (tid <= nb)
		{	// survivors / windows in front of the first record of chain `tid` (tid == nb: all)
			const u32 i = min (s_rs[tid], nrec);
			u32 v;
			if (i >= nrec)
				v = (s_wtot << 12) | s_tot;
			else
			{
				const u32 t = i / RPT;
				v = tab[t];
				for (u32 k = RPT * t; k < i; k++)
					if (flg[k])
						v += ((u32) flg[k] << 12) | 1u;
			}
			s_pos[tid] = v & 0xFFFu;
			s_wpre[tid] = v >> 12;
		}
		__syncthreads ();
		return s_tot;
	};

	// the same for one chain, in any order (a warp reserves its share with one atomic): cheaper, and all a single
	// chain needs.  s_tot / s_wtot must be zero on entry
	auto write_any = [&](u32 nrec, u64 base, u32 slice) -> u32 {
		u32 wsum = 0;
		for (u32 b = 0; b < nrec; b += MG_NT)
		{
			const u32 i = b + tid;
			const u32 f = i < nrec ? flg[i] : 0u;
			const bool keep = f != 0u;
			const u32 bal = __ballot_sync (0xFFFFFFFFu, keep);
			u32 wb = 0;
			if (lane == 0 && bal)
				wb = atomicAdd (&s_tot, (u32) __popc (bal));
			wb = __shfl_sync (0xFFFFFFFFu, wb, 0);
			if (keep)
			{
				const u64 pos = base + wb + __popc (bal & ((1u << lane) - 1u));
				if (pos < mo.out_cap)
				{
					uint4 *dst = reinterpret_cast<uint4 *> (mo.out + pos * RECW);
#pragma unroll
					for (u32 q = 0; q < VEC; q++)
					{
						uint4 x = st4[mg_swz<W> (i, q)];
						if (mo.per_owner && !mo.tagged && q == VEC - 1)
							x.w = slice;
						dst[q] = x;
					}
				}
				wsum += f;
			}
		}
#pragma unroll
		for (int d = 16; d > 0; d >>= 1)
			wsum += __shfl_down_sync (0xFFFFFFFFu, wsum, d);
		if (lane == 0 && wsum)
			atomicAdd (&s_wtot, wsum);
		__syncthreads ();
		return s_tot;
	};

	// One GPU (no owners): a CTA takes a CONTIGUOUS range of chains and writes their survivors one behind the other
	// into a region it reserves once (by its chains' record count), so that a work item can run across groups: the
	// open item is carried from group to group and closed when the next chain would take it over the budget.
	// Several GPUs, sending side: the chains of every owner in pieces of MG_PIECE, dealt to the CTAs round robin;
	// every group's survivors go into its owner's region.
	// The chains' record counts and block lists of the NEXT group are loaded while this one is merged.
	const bool contig = !mo.per_owner;
	auto piece = [&](u32 p, u32 &owner, u32 &lo, u32 &hi) {
		owner = p / ppo;
		lo = min (ch.n_chains, owner * span + (p - owner * ppo) * MG_PIECE);
		hi = min (min (lo + MG_PIECE, owner * span + span), ch.n_chains);
	};
	// one GPU: the CTA's chains [c_lo, c_hi); a group is as many consecutive chains as fit a chunk (found by warp 0
	// when the chains' record counts arrive, a group ahead), at most MG_GMAX
	const u32 cpc = (ch.n_chains + gridDim.x - 1) / gridDim.x;
	const u32 c_lo = (u32) min ((u64) ch.n_chains, (u64) blockIdx.x * cpc), c_hi = (u32) min ((u64) ch.n_chains, (u64) c_lo + cpc);
	u64 cta_base = 0, cta_out = 0;	// (uniform: every thread keeps the same values)
	u64 open_r0 = 0;		// thread 0: the open work item (first record, records, windows)
	u32 open_n = 0, open_w = 0;
	u32 n_staged = 0;		// thread 0: items in s_items[]
	auto flush_items = [&]() {	// (one round trip to the list's counter per MG_ITEMS items instead of one per item)
		if (!n_staged)
			return;
		const u64 ib = atomicAdd (mo.n_items, (unsigned long long) n_staged);
		for (u32 k = 0; k < n_staged; k++)
			if (ib + k < mo.max_items)
				mo.items[ib + k] = s_items[k];
		n_staged = 0;
	};
	auto close_item = [&]() {
		if (!open_n)
			return;
		SkmWork it;
		it.r0 = open_r0;
		it.nrec = open_n;
		it.wsum = open_w;
		it.r = 0;
		it.R = open_w > mo.oversize ? 0u : 1u;
		s_items[n_staged++] = it;
		if (n_staged == MG_ITEMS)
			flush_items ();
		open_n = open_w = 0;
	};
	auto add_chain = [&](u64 r0, u32 nrec, u32 w) {	// thread 0: a chain's survivors (they sit right behind the open item's)
		if (!nrec)
			return;
		if (open_n && open_w + w > mo.budget)
			close_item ();
		if (!open_n)
			open_r0 = r0;
		open_n += nrec;
		open_w += w;
		if (open_w > mo.budget)
			close_item ();	// a single chain beyond the budget: an item of its own
	};
	if (contig && c_lo < c_hi)
	{	// this CTA's chains hold how many records?  (its region in out[]: an upper bound of what survives)
		unsigned long long sum = 0;
		for (u32 c = c_lo + tid; c < c_hi; c += MG_NT)
			sum += (u64) ch.bcount[c] * CH_BLK + min ((u32) ch.head[c], CH_BLK);
#pragma unroll
		for (int d = 16; d > 0; d >>= 1)
			sum += __shfl_down_sync (0xFFFFFFFFu, sum, d);
		if (tid == 0)
			s_start = 0;
		__syncthreads ();
		if (lane == 0 && sum)
			atomicAdd (&s_start, sum);
		__syncthreads ();
		if (tid == 0)
			s_ibase = atomicAdd (mo.out_cursor, s_start);
		__syncthreads ();
		cta_base = s_ibase;
	}
	u32 pb = 0;	// which half of s_nraw / s_b0 holds this group's
	// warp 0: the record counts of up to 32 chains (lane l: chain l of the window, 0 beyond it) -> chains of the
	// group and their records: the longest prefix that fits a chunk (at least one chain).
	auto plan = [&](u32 n, u32 wn, u32 half) {
		u32 incl = min (n, 0x4000000u);	// (only "fits a chunk" matters beyond that)
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const u32 y = __shfl_up_sync (0xFFFFFFFFu, incl, d);
			if (lane >= (u32) d)
				incl += y;
		}
		const u32 k = min (wn, max (1u, (u32) __popc (__ballot_sync (0xFFFFFFFFu, lane < wn && incl <= CH))));
		const u32 tot = __shfl_sync (0xFFFFFFFFu, incl, k ? k - 1 : 0);
		if (lane == 0)
		{
			s_nch[half] = k;
			s_traw[half] = k ? tot : 0u;
		}
	};
	u32 pc = blockIdx.x, c0 = c_lo, p_hi = c_hi, owner = 0;	// the piece, the next chain, the piece's end
	bool more = contig ? c_lo < c_hi : pc < n_pieces;
	if (more)
	{
		if (!contig)
			piece (pc, owner, c0, p_hi);
		const u32 wn = min (MG_GMAX, p_hi - c0);
		u32 n = 0;
		if (tid < wn)
		{
			const u32 c = c0 + tid;
			n = ch.bcount[c] * CH_BLK + min ((u32) ch.head[c], CH_BLK);	// (a chain of 2^32 records or more is not supported)
			s_nraw[tid] = n;
			s_b0[tid] = boff[c];
		}
		if (wid == 0)
			plan (n, wn, 0);
	}
	__syncthreads ();
	for (; more; pb ^= MG_GMAX)
	{
		const u32 nch = s_nch[pb ? 1 : 0];
		// the next group: where it starts is known now; its chains' counts travel while this group is merged
		u32 nx_c0 = c0 + nch, nx_owner = owner, nx_hi = p_hi, nx_pc = pc;
		bool nx_more = true;
		if (nx_c0 >= p_hi)
		{	// the piece is through
			nx_pc = pc + gridDim.x;
			nx_more = !contig && nx_pc < n_pieces;
			if (nx_more)
				piece (nx_pc, nx_owner, nx_c0, nx_hi);
		}
		const u32 nx_wn = nx_more ? min (MG_GMAX, nx_hi - nx_c0) : 0u;
		u32 nx_raw = 0;
		u64 nx_b0 = 0;
		if (tid < nx_wn)
		{
			const u32 c = nx_c0 + tid;
			nx_raw = ch.bcount[c] * CH_BLK + min ((u32) ch.head[c], CH_BLK);
			nx_b0 = boff[c];
		}
		s_nraw = s_nraw2 + pb;
		s_b0 = s_b02 + pb;
		do
		{
		const u32 total_raw = s_traw[pb ? 1 : 0];
		if (total_raw == 0)
			break;
		if (total_raw <= CH && nch <= MG_GMAX)
		{	// ---- the rule: the whole group in one chunk
			{	// first record and first block of every chain in the chunk (warp 0, a scan); the table is cleared meanwhile
				if (wid == 0)
				{
					const u32 n = lane < nch ? s_nraw[lane] : 0u;
					u32 incl = (((n + CH_BLK - 1) / CH_BLK) << 16) | n;	// blocks << 16 | records (<= CH < 2^16)
#pragma unroll
					for (int d = 1; d < 32; d <<= 1)
					{
						const u32 y = __shfl_up_sync (0xFFFFFFFFu, incl, d);
						if (lane >= (u32) d)
							incl += y;
					}
					const u32 excl = incl - ((((n + CH_BLK - 1) / CH_BLK) << 16) | n);
					if (lane < nch)
					{
						s_rs[lane] = excl & 0xFFFFu;
						s_bs[lane] = excl >> 16;
					}
					if (lane == nch - 1)
					{
						s_rs[nch] = incl & 0xFFFFu;
						s_bs[nch] = incl >> 16;
					}
				}
				for (u32 v = tid; v < TS; v += MG_NT)
					tab[v] = 0u;
				__syncthreads ();
				// the chunk's blocks: where each one is, where its records go, how many it holds
				const u32 nbt = s_bs[nch];
				if (tid < nbt)
				{
					u32 g = 0;
					while (s_bs[g + 1] <= tid)
						g++;
					const u32 j = tid - s_bs[g];
					sbl[tid] = j == 0 ? c0 + g : __ldg (blist + s_b0[g] + j - 1);
					sbd[tid] = (s_rs[g] + CH_BLK * j) | (min (CH_BLK, s_nraw[g] - CH_BLK * j) << 16);
				}
				__syncthreads ();
				// the records: one flat loop over (block, record, 16-byte part), four loads in flight per thread
				constexpr u32 SLOTS = CH_BLK * VEC;
				const u32 tot = nbt * SLOTS;
				uint4 *dst = st4;
				for (u32 v0 = tid; v0 < tot; v0 += 4 * MG_NT)
				{
					uint4 x[4];
					u32 di[4];
#pragma unroll
					for (u32 k = 0; k < 4; k++)
					{
						const u32 v = v0 + k * MG_NT;
						di[k] = 0xFFFFFFFFu;
						if (v < tot)
						{
							const u32 b = v / SLOTS, w = v - b * SLOTS, i = w / VEC, part = w - i * VEC;
							const u32 d = sbd[b];
							if (i < (d >> 16))
							{
								x[k] = ldg_stream (reinterpret_cast<const uint4 *> (ch.recs + ((u64) sbl[b] * CH_BLK + i) * RECW) + part);
								if (!HAS_MULT && part == 0)
									x[k].z = 1u;
								di[k] = mg_swz<W> ((d & 0xFFFFu) + i, part);
							}
						}
					}
#pragma unroll
					for (u32 k = 0; k < 4; k++)
						if (di[k] != 0xFFFFFFFFu)
							dst[di[k]] = x[k];
				}
			}
			__syncthreads ();
			dedupe (total_raw);
			__syncthreads ();
			u64 base = cta_base + cta_out;
			if (mo.per_owner)
			{	// exact space in the owner's region: count the survivors first
				u32 k = 0;
				for (u32 i = tid; i < total_raw; i += MG_NT)
					k += flg[i] != 0;
#pragma unroll
				for (int d = 16; d > 0; d >>= 1)
					k += __shfl_down_sync (0xFFFFFFFFu, k, d);
				if (tid == 0)
					s_tot = 0;
				__syncthreads ();
				if (lane == 0 && k)
					atomicAdd (&s_tot, k);
				__syncthreads ();
				if (tid == 0)
					s_start = mo.region[owner] + atomicAdd (mo.rcur + owner, (unsigned long long) s_tot);
				__syncthreads ();
				base = s_start;
			}
			u32 kept;
			if (nch == 1)
			{
				if (tid == 0)
					s_tot = s_wtot = 0;
				__syncthreads ();
				kept = write_any (total_raw, base, c0);
				if (tid == 0)
				{
					s_pos[0] = s_wpre[0] = 0;
					s_pos[1] = kept;
					s_wpre[1] = s_wtot;
				}
			}
			else
				kept = write_out (total_raw, base, nch, c0);
			if (tid == 0)
			{
				kept_total += kept;
				win_total += s_wtot;
				if (contig)	// work items: the chains in order, each behind the other
					for (u32 g = 0; g < nch; g++)
						add_chain (base + s_pos[g], s_pos[g + 1] - s_pos[g], s_wpre[g + 1] - s_wpre[g]);
			}
			cta_out += kept;
			__syncthreads ();
			break;
		}
		// ---- a group with more records than a chunk holds: chain by chain, a chain in as many chunks as it takes.
		// Every chain is an item of its own (copies that sit in different chunks of a chain are not merged).
		for (u32 g = 0; g < nch; g++)
		{
			const u32 n_raw = s_nraw[g];
			const u64 b0 = s_b0[g];
			const u64 first = cta_out, item_base = cta_base + cta_out;
			u32 wsum = 0;
			for (u32 q0 = 0; q0 < n_raw; q0 += CH)
			{
				const u32 nrec = min (CH, n_raw - q0);
				const u32 nbk = (nrec + CH_BLK - 1) / CH_BLK, lb0 = q0 / CH_BLK;
				for (u32 j = tid; j < nbk; j += MG_NT)
					sbl[j] = lb0 + j == 0 ? c0 + g : blist[b0 + lb0 + j - 1];
				for (u32 v = tid; v < TS; v += MG_NT)
					tab[v] = 0u;
				if (tid == 0)
				{
					s_rs[0] = 0;
					s_rs[1] = nrec;
				}
				__syncthreads ();
				for (u32 v = tid; v < nrec * VEC; v += MG_NT)
				{
					const u32 i = v / VEC, part = v - i * VEC;
					uint4 x = ldg_stream (reinterpret_cast<const uint4 *> (ch.recs + ((u64) sbl[i / CH_BLK] * CH_BLK + (i & (CH_BLK - 1))) * RECW) + part);
					if (!HAS_MULT && part == 0)
						x.z = 1u;
					st4[mg_swz<W> (i, part)] = x;
				}
				__syncthreads ();
				dedupe (nrec);
				__syncthreads ();
				u64 base;
				if (mo.per_owner)
				{
					u32 k = 0;
					for (u32 i = tid; i < nrec; i += MG_NT)
						k += flg[i] != 0;
#pragma unroll
					for (int d = 16; d > 0; d >>= 1)
						k += __shfl_down_sync (0xFFFFFFFFu, k, d);
					if (tid == 0)
						s_tot = 0;
					__syncthreads ();
					if (lane == 0 && k)
						atomicAdd (&s_tot, k);
					__syncthreads ();
					if (tid == 0)
						s_start = mo.region[owner] + atomicAdd (mo.rcur + owner, (unsigned long long) s_tot);
					__syncthreads ();
					base = s_start;
				}
				else
					base = cta_base + cta_out;
				if (tid == 0)
					s_tot = s_wtot = 0;
				__syncthreads ();
				const u32 kept = write_any (nrec, base, c0 + g);
				cta_out += kept;
				wsum += s_wtot;
				__syncthreads ();
			}
			if (tid == 0)
			{
				const u64 nrec = cta_out - first;
				kept_total += nrec;
				win_total += wsum;
				if (contig)
				{	// a chain that took several chunks is an item of its own (copies in different chunks were not merged,
					// but they sit in one item)
					if (n_raw > CH)
						close_item ();
					add_chain (item_base, (u32) min (nrec, (u64) 0xFFFFFFFFu), wsum);
					if (n_raw > CH)
						close_item ();
				}
			}
		}
		__syncthreads ();
		} while (0);
		// the next group's chains
		if (tid < MG_GMAX)
		{
			s_nraw2[(pb ^ MG_GMAX) + tid] = nx_raw;
			s_b02[(pb ^ MG_GMAX) + tid] = nx_b0;
			plan (nx_raw, nx_wn, pb ? 0 : 1);
		}
		__syncthreads ();
		c0 = nx_c0;
		owner = nx_owner;
		p_hi = nx_hi;
		pc = nx_pc;
		more = nx_more;
	}
	if (tid == 0)
	{
		close_item ();
		flush_items ();
	}
	if (tid == 0 && kept_total)
	{
		atomicAdd (mo.n_kept, kept_total);
		atomicAdd (mo.n_kept + 1, win_total);
	}
}

// records per owner rank (multi-GPU send side): sum of the chains' record counts over each rank's range of slices
__global__ void chain_owner_count_kernel (SkmChains ch, u32 per_owner, unsigned long long *counts)
{
	__shared__ unsigned long long s_sum;
	for (u32 owner = blockIdx.x; owner * (u64) per_owner < ch.n_chains; owner += gridDim.x)
	{
		if (threadIdx.x == 0)
			s_sum = 0;
		__syncthreads ();
		unsigned long long s = 0;
		const u32 lo = owner * per_owner, hi = min (ch.n_chains, lo + per_owner);
		for (u32 c = lo + threadIdx.x; c < hi; c += blockDim.x)
			s += (u64) ch.bcount[c] * CH_BLK + min ((u32) ch.head[c], CH_BLK);
#pragma unroll
		for (int d = 16; d > 0; d >>= 1)
			s += __shfl_down_sync (0xFFFFFFFFu, s, d);
		if ((threadIdx.x & 31) == 0 && s)
			atomicAdd (&s_sum, s);
		__syncthreads ();
		if (threadIdx.x == 0)
			counts[owner] = s_sum;
		__syncthreads ();
	}
}

// ------------------------------------------------------------------------------------------------
// Records that are already made -> chains.  The receiving side of the multi-GPU exchange: the slice
// travels in the record's last word; this rank owns slices [lo, lo + n_local).
static constexpr int AP_NT = 256, AP_TILE = 2048;	// records a CTA appends between two looks at its block range

template <int RECW>
__global__ void __launch_bounds__ (AP_NT)
skm_append_kernel (SkmChains ch, const u32 *rec, u64 n, u32 lo)
{
	__shared__ u32 s_pool[2];
	pool_begin (ch, s_pool);
	__syncthreads ();
	const u64 n_tiles = (n + AP_TILE - 1) / AP_TILE;
	for (u64 t = blockIdx.x; t < n_tiles; t += gridDim.x)
	{
		if (threadIdx.x == 0)
			pool_refill (ch, s_pool, 128);
		__syncthreads ();
		// A thread takes AP_TILE / AP_NT CONSECUTIVE records: the records of a slice arrive next to each other (a
		// sender's run of a few dozen), and threads that append to the same chain at the same time wait for each
		// other at every block boundary (with one record per thread and turn, 16 threads met in a chain: the
		// pass took four times as long on 8 GPUs as on 2).  The next record's position is drawn while this one is written.
		if (s_pool[1] != 0)
		{
			constexpr u32 PER = AP_TILE / AP_NT, VEC = RECW / 4;
			const u64 i0 = t * AP_TILE + (u64) threadIdx.x * PER, i1 = min (n, i0 + PER);
			uint4 nx[VEC];
			u32 s_nx = 0;
			u64 tk_nx = 0;
			bool ok_nx = false;
			auto fetch = [&](u64 i) {
				const uint4 *src = reinterpret_cast<const uint4 *> (rec + i * RECW);
#pragma unroll
				for (u32 q = 0; q < VEC; q++)
					nx[q] = ldg_stream (src + q);
				s_nx = nx[VEC - 1].w - lo;
				nx[VEC - 1].w = 0;
				ok_nx = s_nx < ch.n_chains;
				if (ok_nx)
					tk_nx = chain_ticket (ch, s_nx);
				else	// not this rank's: the exchange went wrong
					atomicOr (reinterpret_cast<unsigned long long *> (&ch.ctr->overflow), (unsigned long long) OVF_FOREIGN);
			};
			if (i0 < i1)
				fetch (i0);
			for (u64 i = i0; i < i1; i++)
			{
				uint4 v[VEC];
#pragma unroll
				for (u32 q = 0; q < VEC; q++)
					v[q] = nx[q];
				const u32 sl = s_nx;
				const u64 tk = tk_nx;
				const bool ok = ok_nx;
				ok_nx = false;
				if (i + 1 < i1)
					fetch (i + 1);
				if (!ok)
					continue;
				u32 *dst = chain_place (ch, sl, tk, s_pool, ok_nx ? &tk_nx : nullptr, s_nx);
				if (dst)
				{
#pragma unroll
					for (u32 q = 0; q < VEC; q++)
						reinterpret_cast<uint4 *> (dst)[q] = v[q];
				}
			}
		}
		__syncthreads ();
	}
	pool_end (ch, s_pool);
}

// ------------------------------------------------------------------------------------------------
// A slice that overflowed its image by far is cut into q sub-slices by k-mer hash in ONE pass over
// its records: every window becomes a one-window sub-record (the window's K bases and its
// neighbours, the record's multiplicity, the window's ordinal) in the chain of sub-slice qbase +
// bucket.  The sub-slices then go through skm_merge_kernel and skm_build_kernel like any slice.
// (Retrying a slice as q work items that each filter the windows by hash scans its records q times;
// the slices that overflow by far are the highly expressed loci — most of the reads of a skewed data set.)
struct SkmSplit { u64 r0; u32 n, q, qbase, pad; };	// records r0 .. r0 + n of the merged runs -> sub-slices qbase .. qbase + q
static constexpr int SP_NT = 256, SP_RECS = 256;	// records per piece of a split (one look at the CTA's block range per piece)

template <int W>
__global__ void __launch_bounds__ (SP_NT)
skm_split_kernel (SkmChains sub, const u32 *rec2, const SkmSplit *splits, u32 n_splits, int K)
{
	constexpr u32 RECW = SkmRec<W>::WORDS, LAST = RECW - SKM_HDR - 1;
	__shared__ u32 s_pool[2];
	pool_begin (sub, s_pool);
	__syncthreads ();
	Key<W> kmask;
#pragma unroll
	for (int q = 0; q < W; q++)
	{
		const int bits = 2 * K - 64 * (W - 1 - q);
		kmask.w[q] = bits >= 64 ? ~0ull : (bits > 0 ? (1ull << bits) - 1 : 0ull);
	}
	const int top = 2 * (K - 1);
	for (u32 sp = blockIdx.x; sp < n_splits; sp += gridDim.x)
	{
		const SkmSplit S = splits[sp];
		const u64 r1 = S.r0 + S.n;
		// one record per thread and step; the CTA's block range is looked at once per step
		for (u64 base = S.r0; base < r1; base += SP_NT)
		{
			if (threadIdx.x == 0)	// SP_NT records x ~8 windows each: ~64 blocks fill up per step, 16x that when every record is full
				pool_refill (sub, s_pool, 256);
			__syncthreads ();
			const u64 i = base + threadIdx.x;
			if (i < r1 && s_pool[1] != 0)
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
					u32 left, right, ha, hb;
					skm_roll_window<W> (st, key, left, right);
					skm_hash2<W> (key, ha, hb);
					const u32 g = S.qbase + __umulhi (hb * 0x9E3779B1u, S.q);
					u32 *dst = chain_append (sub, g, s_pool);
					if (dst)
					{
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
			__syncthreads ();
		}
	}
	pool_end (sub, s_pool);
}

}	// namespace sdt
