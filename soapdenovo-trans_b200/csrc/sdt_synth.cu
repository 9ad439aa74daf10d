// sdt_synth.cu — synthetic paired-end reads generated directly in device memory, bit-identical to
// soapdenovo-trans_b200/synth.py (make_reads + pack_reads).  Bench/test utility: lets the full-size
// BASELINE.json configs (50-100 M reads) exist in HBM without ever touching a file or the host.
#include "../../include/sdtgpu.h"
#include <cuda_runtime.h>
#include <cstdint>

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

__device__ __forceinline__ u64 mix64 (u64 x)
{
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}
__device__ __forceinline__ u64 h3 (u64 seed, u64 a, u64 b)
{
	return mix64 (mix64 (seed + a * 0x9E3779B97F4A7C15ull) ^ (b * 0xD1B54A32D192ED03ull));
}

__global__ void __launch_bounds__ (256)
synth_kernel (const uint8_t *tr, const u64 *starts, const u32 *lengths, const u64 *cum, u32 T, u64 seed,
	      u64 first_pair, u64 n_pairs, u32 L, u32 stride_words, u32 *out)
{
	const u64 n_words = 2 * n_pairs * stride_words;
	for (u64 idx = blockIdx.x * 256ull + threadIdx.x; idx < n_words; idx += (u64) gridDim.x * 256ull)
	{
		const u64 rl = idx / stride_words;
		const u32 q = (u32) (idx - rl * stride_words);
		u32 word = 0;
		if (16 * q < L)
		{
			const u64 p = first_pair + (rl >> 1);
			const u32 mate = (u32) (rl & 1);
			const u64 rg = 2 * first_pair + rl;
			const u64 u = h3 (seed, p, 0) >> 24;	// 40-bit draw against the cumulative weights
			u32 lo = 0, hi = T - 1;			// first t with cum[t] > u
			while (lo < hi)
			{
				const u32 mid = (lo + hi) >> 1;
				if (cum[mid] > u)
					hi = mid;
				else
					lo = mid + 1;
			}
			const u64 tlen = lengths[lo];
			const u64 hf = h3 (seed, p, 1);
			u64 s = 0;
#pragma unroll
			for (int k = 0; k < 8; k++)
				s += (hf >> (8 * k)) & 0xFF;
			u64 frag = (20 * s + 21400) / 209;
			if (frag < L)
				frag = L;
			if (frag > tlen)
				frag = tlen;
			const u64 start = h3 (seed, p, 2) % (tlen - frag + 1);
			const u64 base0 = starts[lo] + start;
			for (u32 i = 0; i < 16; i++)
			{
				const u32 b = 16 * q + i;
				u32 code = 0;
				if (b < L)
				{
					code = mate ? (tr[base0 + frag - 1 - b] ^ 2u) : tr[base0 + b];
					const u64 e = h3 (seed ^ 0xEull, rg, b);
					if ((e & 0xFFFF) < 655)
						code = (code + 1 + (u32) ((e >> 16) % 3)) & 3u;
				}
				// tight string: base i of this word sits in byte i/4 at bits 7-2*(i%4)..; little-endian u32 store
				word |= code << (8 * (i >> 2) + 6 - 2 * (i & 3));
			}
		}
		out[idx] = word;
	}
}

}	// namespace

extern "C" int sdtgpu_synth_reads_device (int device, void *stream, const uint8_t *d_tr_bases, const uint64_t *d_starts,
					  const uint32_t *d_lengths, const uint64_t *d_cum, uint32_t n_transcripts,
					  uint64_t seed, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len,
					  uint32_t stride_bytes, uint8_t *d_packed_out)
{
	if (!d_tr_bases || !d_starts || !d_lengths || !d_cum || !d_packed_out || n_transcripts == 0 ||
	    (stride_bytes & 3) || stride_bytes * 4 < read_len || ((uintptr_t) d_packed_out & 3))
		return SDTGPU_EINVAL;
	if (n_pairs == 0)
		return SDTGPU_OK;
	if (cudaSetDevice (device) != cudaSuccess)
		return SDTGPU_ECUDA;
	const u64 n_words = 2 * n_pairs * (stride_bytes / 4);
	const unsigned grid = (unsigned) (n_words / 256 + 1 < 148ull * 32 ? n_words / 256 + 1 : 148ull * 32);
	synth_kernel<<<grid, 256, 0, (cudaStream_t) stream>>> (d_tr_bases, (const u64 *) d_starts, d_lengths, (const u64 *) d_cum,
							       n_transcripts, seed, first_pair, n_pairs, read_len, stride_bytes / 4,
							       reinterpret_cast<u32 *> (d_packed_out));
	return cudaGetLastError () == cudaSuccess ? SDTGPU_OK : SDTGPU_ECUDA;
}
