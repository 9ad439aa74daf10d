// ldvar_bench.cu — how many DRAM bytes does one random 16/32-byte access cost on B200, per load flavour?
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ u64 mix(u64 k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }

template <int MODE>
__global__ void __launch_bounds__(256) k(u64* tab, u64 cap, u64 n, u64 seed, u64* sink, u64 policy)
{
	u64 acc = 0;
	for (u64 i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (u64)gridDim.x * 256ull) {
		u64* p = tab + 4 * __umul64hi(mix(seed + i), cap);
		u64 a = 0, b = 0;
		if (MODE == 0) asm volatile("ld.global.relaxed.gpu.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		if (MODE == 1) asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		if (MODE == 2) asm volatile("ld.global.cs.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		if (MODE == 3) asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		if (MODE == 4) asm volatile("ld.global.relaxed.sys.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		if (MODE == 5) asm volatile("ld.global.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(a), "=l"(b) : "l"(p), "l"(policy) : "memory");
		if (MODE == 6) asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		if (MODE == 7) { a = atomicOr(p, 0ull); }						// atomic read 8 B
		if (MODE == 8) asm volatile("{\n\t.reg .b128 c, n, o;\n\tmov.b128 c, {%2,%3};\n\tmov.b128 n, {%2,%3};\n\tatom.global.relaxed.gpu.cas.b128 o, [%4], c, n;\n\tmov.b128 {%0,%1}, o;\n\t}" : "=l"(a), "=l"(b) : "l"(0x5555555555555555ull), "l"(0x3333333333333333ull), "l"(p) : "memory");	// atomic read 16 B
		if (MODE == 9) asm volatile("ld.global.relaxed.gpu.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
		if (MODE == 10) asm volatile("ld.global.relaxed.gpu.u32 %0, [%1];" : "=r"(*(u32*)&a) : "l"(p) : "memory");
		if (MODE == 11) asm volatile("ld.global.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		if (MODE == 12) asm volatile("ld.global.lu.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
		acc += a + b;
	}
	if (acc == 0x1234567) *sink = acc;
}
template <int MODE> void run(const char* name, u64* tab, u64 cap, u64 n, u64* sink, u64 policy)
{
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	float best = 1e30f;
	for (int it = 0; it < 3; it++) {
		CK(cudaEventRecord(e0));
		k<MODE><<<148 * 8, 256>>>(tab, cap, n, 777ull * (it + 1), sink, policy);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it && ms < best) best = ms;
	}
	printf("MODE %2d %-40s %8.3f ms %7.2f G/s\n", MODE, name, best, n / best * 1e-6); fflush(stdout);
}
__global__ void mkpolicy(u64* out) { u64 p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); *out = p; }
int main(int argc, char** argv)
{
	int gran = argc > 1 ? atoi(argv[1]) : 0;
	if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
	const u64 cap = 1ull << 29, n = 1ull << 27;
	u64* tab; CK(cudaMalloc(&tab, cap * 32)); CK(cudaMemset(tab, 0, cap * 32));
	u64* sink; CK(cudaMalloc(&sink, 16));
	mkpolicy<<<1, 1>>>(sink + 1); u64 policy; CK(cudaMemcpy(&policy, sink + 1, 8, cudaMemcpyDeviceToHost));
	run<0>("ld.relaxed.gpu.v2.u64", tab, cap, n, sink, policy);
	run<1>("ld.cv.v2.u64", tab, cap, n, sink, policy);
	run<2>("ld.cs.v2.u64", tab, cap, n, sink, policy);
	run<3>("ld.volatile.v2.u64", tab, cap, n, sink, policy);
	run<4>("ld.relaxed.sys.v2.u64", tab, cap, n, sink, policy);
	run<5>("ld.L2::cache_hint evict_first", tab, cap, n, sink, policy);
	run<6>("ld.nc.L1::no_allocate", tab, cap, n, sink, policy);
	run<7>("atomicOr(p,0) 8 B", tab, cap, n, sink, policy);
	run<8>("atom.cas.b128 (never matches) 16 B", tab, cap, n, sink, policy);
	run<9>("ld.relaxed.gpu.u64", tab, cap, n, sink, policy);
	run<10>("ld.relaxed.gpu.u32", tab, cap, n, sink, policy);
	run<11>("ld.L2::64B.v2.u64", tab, cap, n, sink, policy);
	run<12>("ld.lu.v2.u64", tab, cap, n, sink, policy);
	return 0;
}
