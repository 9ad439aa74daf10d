// pagecoh_bench.cu — is B200's random-access cap an address-translation limit?
// Every access is to a cold 32-byte slot of a 16 GiB table, but the lanes of a warp (or `G` consecutive
// lanes) pick their slots inside ONE randomly chosen region of R bytes.  If the 36.65 G requests/s
// cap comes from TLB misses, confining a warp to one 2 MiB page must lift it; if it comes from the
// L2/DRAM side, it must not.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ u64 mix(u64 k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }
template <int MODE> __global__ void __launch_bounds__(256) k(u64* tab, u64 cap, u64 n, u64 seed, u64* sink, u64 region_slots, u32 group)
{
	u64 acc = 0;
	for (u64 i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (u64)gridDim.x * 256ull) {
		const u64 g = i / group;                                   // lanes of one group share a region
		const u64 region = __umul64hi(mix(seed ^ (g * 0x9E3779B97F4A7C15ull)), cap / region_slots);
		u64* p = tab + 4 * (region * region_slots + __umul64hi(mix(seed + i), region_slots));
		u64 a = 0, b = 0;
		asm volatile("ld.global.relaxed.gpu.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		if (MODE == 1) a += atomicCAS(p + 1, b, b + 1);
		acc += a + b;
	}
	if (acc == 0x1234567) *sink = acc;
}
int main()
{
	const u64 cap = 1ull << 29, n = 1ull << 28;
	u64* tab; CK(cudaMalloc(&tab, cap * 32)); CK(cudaMemset(tab, 0, cap * 32));
	u64* sink; CK(cudaMalloc(&sink, 8));
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	const u64 regions[] = { 4096 / 32, 65536 / 32, (2ull << 20) / 32, (32ull << 20) / 32, (512ull << 20) / 32, cap };
	const u32 groups[] = { 1, 8, 32, 256, 4096 };
	for (u64 rs : regions) for (u32 g : groups) for (int mode = 0; mode < 2; mode++) {
		float best = 1e30f;
		for (int it = 0; it < 3; it++) {
			CK(cudaEventRecord(e0));
			if (mode == 0) k<0><<<148 * 8, 256>>>(tab, cap, n, 31337ull * (it + 1), sink, rs, g); else k<1><<<148 * 8, 256>>>(tab, cap, n, 31337ull * (it + 1), sink, rs, g);
			CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
			float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it && ms < best) best = ms;
		}
		printf("region %10llu B  group %5u lanes  %-9s %7.2f G/s\n", rs * 32, g, mode ? "ld+CAS64" : "ld128", n / best * 1e-6); fflush(stdout);
	}
	return 0;
}
