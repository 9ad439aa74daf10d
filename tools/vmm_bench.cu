// vmm_bench.cu — does the page size behind the table change B200's random-access rate?
// Allocates the table (a) with cudaMalloc, (b) through the VMM API (cuMemCreate/cuMemMap) with the
// minimum and the recommended granularity, and measures random 16-byte loads and atomics.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e, &s); printf("%s: %s\n", #x, s); exit(1); } } while (0)
__device__ __forceinline__ u64 mix(u64 k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }
template <int MODE> __global__ void __launch_bounds__(256) k(u64* tab, u64 cap, u64 n, u64 seed, u64* sink)
{
	u64 acc = 0;
	for (u64 i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (u64)gridDim.x * 256ull) {
		u64* p = tab + 4 * __umul64hi(mix(seed + i), cap);
		u64 a = 0, b = 0;
		if (MODE == 0) asm volatile("ld.global.relaxed.gpu.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		if (MODE == 1) { asm volatile("ld.global.relaxed.gpu.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); a = atomicCAS(p + 1, b, b + 1); }
		acc += a + b;
	}
	if (acc == 0x1234567) *sink = acc;
}
static void bench(const char* what, u64* tab, u64 cap, u64* sink)
{
	const u64 n = 1ull << 28;
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int mode = 0; mode < 2; mode++) {
		float best = 1e30f;
		for (int it = 0; it < 3; it++) {
			CK(cudaEventRecord(e0));
			if (mode == 0) k<0><<<148 * 8, 256>>>(tab, cap, n, 99 * (it + 1), sink); else k<1><<<148 * 8, 256>>>(tab, cap, n, 99 * (it + 1), sink);
			CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
			float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it && ms < best) best = ms;
		}
		printf("%-44s %-12s %7.2f G/s\n", what, mode ? "ld+CAS64" : "ld128", n / best * 1e-6);
	}
}
int main(int argc, char** argv)
{
	CK(cudaFree(0));
	const size_t bytes = (argc > 1 ? strtoull(argv[1], 0, 10) : 16ull) << 30; const u64 cap = bytes / 32;
	u64* sink; CK(cudaMalloc(&sink, 8));
	{ u64* t; CK(cudaMalloc(&t, bytes)); CK(cudaMemset(t, 0, bytes)); bench("cudaMalloc", t, cap, sink); CK(cudaFree(t)); }
	CUmemAllocationProp prop = {}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
	size_t gmin = 0, grec = 0;
	CU(cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
	CU(cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
	printf("VMM granularity: minimum %zu B, recommended %zu B\n", gmin, grec);
	size_t grans[3] = { gmin, grec, 512ull << 20 };
	for (int gi = 0; gi < 3; gi++) {
		const size_t g = grans[gi];
		CUdeviceptr va = 0; CU(cuMemAddressReserve(&va, bytes, g, 0, 0));
		// one physical handle per `g` bytes vs one handle for everything
		for (int whole = 0; whole < 2; whole++) {
			const size_t chunk = whole ? bytes : g; const size_t nchunks = bytes / chunk;
			CUmemGenericAllocationHandle* hs = (CUmemGenericAllocationHandle*)malloc(sizeof(*hs) * nchunks);
			bool ok = true;
			for (size_t c = 0; c < nchunks && ok; c++) { ok = cuMemCreate(&hs[c], chunk, &prop, 0) == CUDA_SUCCESS; if (ok) CU(cuMemMap(va + c * chunk, chunk, 0, hs[c], 0)); }
			if (!ok) { printf("cuMemCreate failed for chunk %zu\n", chunk); continue; }
			CUmemAccessDesc ad = {}; ad.location = prop.location; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
			CU(cuMemSetAccess(va, bytes, &ad, 1));
			CK(cudaMemset((void*)va, 0, bytes));
			char name[128]; snprintf(name, sizeof name, "VMM align %zu MiB, handles of %zu MiB", g >> 20, chunk >> 20);
			bench(name, (u64*)va, cap, sink);
			CU(cuMemUnmap(va, bytes));
			for (size_t c = 0; c < nchunks; c++) CU(cuMemRelease(hs[c]));
			free(hs);
		}
		CU(cuMemAddressFree(va, bytes));
	}
	return 0;
}
