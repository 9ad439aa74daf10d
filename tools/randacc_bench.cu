// randacc_bench.cu — measures what B200's memory system sustains for the access pattern of a hash
// table upsert: random 32-byte-slot loads / REDs / CASes over a table far larger than L2.
// This is the "random-access roofline" the insert kernel is judged against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o randacc_bench randacc_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ u64 mix(u64 k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }
__device__ __forceinline__ void ld256(const void* p, u64& a, u64& b, u64& c, u64& d) { asm volatile("ld.global.relaxed.gpu.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory"); }
__device__ __forceinline__ void ld128(const void* p, u64& a, u64& b) { asm volatile("ld.global.relaxed.gpu.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ void ld128cg(const void* p, u64& a, u64& b) { asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ void ld256cg(const void* p, u64& a, u64& b, u64& c, u64& d) { asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory"); }
__device__ __forceinline__ void red32(u32* p, u32 v) { asm volatile("red.global.relaxed.gpu.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

// windowed variant: record i touches a random slot of region (i / per_region); regions tile the table in order.
// Models the partitioned insert: all records of one slot-range bucket are processed together.
__device__ u64 g_region_slots = 0, g_per_region = 0;
template <int MODE, int U>
__global__ void __launch_bounds__(256) k(u64* tab, u64 cap, u64 n, u64 seed, u64* sink)
{
	const u64 region_slots = g_region_slots, per_region = g_per_region;
	u64 acc = 0;
	for (u64 i0 = (blockIdx.x * 256ull + threadIdx.x) * U; i0 < n; i0 += (u64)gridDim.x * 256ull * U) {
		u64* p[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			if (region_slots) { u64 r = (i0 + u) / per_region; p[u] = tab + 4 * ((r * region_slots) % (cap - region_slots) + __umul64hi(mix(seed + i0 + u), region_slots)); }
			else p[u] = tab + 4 * __umul64hi(mix(seed + i0 + u), cap);
		}
		u64 a[U], b[U], c[U], d[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			a[u] = b[u] = c[u] = d[u] = 0;
			if (MODE == 0) ld256(p[u], a[u], b[u], c[u], d[u]);
			if (MODE == 1) { ld128(p[u], a[u], b[u]); ld128(p[u] + 2, c[u], d[u]); }
			if (MODE == 2 || MODE == 6 || MODE == 7 || MODE == 8 || MODE == 9 || (MODE >= 12 && MODE <= 19)) ld128(p[u], a[u], b[u]);
			if (MODE == 3) ld128cg(p[u], a[u], b[u]);
			if (MODE == 4) ld256cg(p[u], a[u], b[u], c[u], d[u]);
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			if (MODE == 5 || MODE == 6 || MODE == 8 || MODE == 9) red32((u32*)(p[u] + 3), 1u);
			if (MODE == 7 || MODE == 8) { u64 old = atomicCAS(p[u] + 1, b[u], b[u] + 1); acc += old; }
			if (MODE == 9) { red32((u32*)(p[u] + 1), 1u); }
			if (MODE == 10) { u64 old = atomicCAS(p[u] + 1, 0ull, 1ull); acc += old; }
			if (MODE == 11) { u64 old = atomicAdd(p[u] + 1, 1ull); acc += old; }
			if (MODE == 12) { u32 old = atomicAdd((u32*)(p[u] + 3), 1u); acc += old; }				// ld + returning add (other half)
			if (MODE == 13) { red32((u32*)(p[u] + 1), 1u); }							// ld + red, same 16-B half
			if (MODE == 14) { asm volatile("st.global.relaxed.gpu.u32 [%0], %1;" :: "l"(p[u] + 3), "r"((u32)b[u] + 1) : "memory"); }	// ld + plain store
			if (MODE == 15) { u64 o1 = atomicCAS(p[u] + 1, b[u], b[u] + 1); u64 o2 = atomicCAS(p[u] + 3, 0ull, o1); acc += o2; }	// ld + 2 dependent CAS
			if (MODE == 16) { u64 o1 = atomicCAS(p[u] + 1, b[u], b[u] + 1); u32 o2 = atomicAdd((u32*)(p[u] + 3), 1u); acc += o1 + o2; }	// ld + CAS + returning add (independent)
			if (MODE == 17) { u32 o2 = atomicAdd((u32*)(p[u] + 3), 1u); u64 o1 = atomicCAS(p[u] + 1, b[u], b[u] + 1); acc += o1 + o2; }
			if (MODE == 18) { u64 o1 = atomicCAS(p[u] + 1, b[u], b[u] + 1); acc += o1; red32((u32*)(p[u] + 3), 1u); }		// CAS first, then red
			if (MODE == 19) { u64 oa, ob; asm volatile("{\n\t.reg .b128 c, n, o;\n\tmov.b128 c, {%2,%3};\n\tmov.b128 n, {%4,%5};\n\tatom.global.relaxed.gpu.cas.b128 o, [%6], c, n;\n\tmov.b128 {%0,%1}, o;\n\t}" : "=l"(oa), "=l"(ob) : "l"(a[u]), "l"(b[u]), "l"(a[u] + 1), "l"(b[u] + 1), "l"(p[u]) : "memory"); acc += oa + ob; }	// ld + CAS128
			acc += a[u] + b[u] + c[u] + d[u];
		}
	}
	if (acc == 0x1234567) *sink = acc;
}

template <int MODE, int U> void run(const char* name, u64* tab, u64 cap, u64 n, u64* sink, int sms)
{
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE, U>, 256, 0));
	float best = 1e30f;
	for (int it = 0; it < 4; it++) {
		CK(cudaEventRecord(e0));
		k<MODE, U><<<sms * occ, 256>>>(tab, cap, n, 1234567ull * (it + 1), sink);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it && ms < best) best = ms;
	}
	printf("%-44s U=%d occ=%d  %8.3f ms  %7.2f G slots/s  (x32B = %7.1f GB/s, x64B = %7.1f GB/s)\n", name, U, occ, best, n / best * 1e-6, n / best * 1e-6 * 32, n / best * 1e-6 * 64);
	fflush(stdout);
}

int main(int argc, char** argv)
{
	int gran = argc > 1 ? atoi(argv[1]) : 0;
	if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
	size_t g = 0; CK(cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity));
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	printf("L2 fetch granularity %zu B, %d SMs\n", g, prop.multiProcessorCount);
	const u64 cap = argc > 3 ? strtoull(argv[3], 0, 10) : (1ull << 29);	// default 512 Mi slots x 32 B = 16 GiB
	const u64 n = argc > 2 ? strtoull(argv[2], 0, 10) : (1ull << 29);
	u64* tab; CK(cudaMalloc(&tab, cap * 32)); CK(cudaMemset(tab, 0, cap * 32));
	u64* sink; CK(cudaMalloc(&sink, 8));
	int sms = prop.multiProcessorCount;
	for (int pass = 0; pass < 3; pass++) {
		u64 rs = pass == 0 ? 0 : (pass == 1 ? (8ull << 20) / 32 : (32ull << 20) / 32), pr = rs * 3;
		CK(cudaMemcpyToSymbol(g_region_slots, &rs, 8)); CK(cudaMemcpyToSymbol(g_per_region, &pr, 8));
		printf("== %s\n", pass == 0 ? "uniform random over the whole table" : (pass == 1 ? "windowed 8 MiB regions" : "windowed 32 MiB regions"));
		run<2, 1>("ld128", tab, cap, n, sink, sms);
		run<5, 1>("red only", tab, cap, n, sink, sms);
		run<10, 1>("CAS64 only", tab, cap, n, sink, sms);
		run<6, 1>("ld128 + red (other half)", tab, cap, n, sink, sms);
		run<13, 1>("ld128 + red (same half)", tab, cap, n, sink, sms);
		run<12, 1>("ld128 + returning add32", tab, cap, n, sink, sms);
		run<14, 1>("ld128 + plain store", tab, cap, n, sink, sms);
		run<7, 1>("ld128 + CAS64", tab, cap, n, sink, sms);
		run<19, 1>("ld128 + CAS128", tab, cap, n, sink, sms);
		run<15, 1>("ld128 + CAS64 + dependent CAS64", tab, cap, n, sink, sms);
		run<16, 1>("ld128 + CAS64 + returning add32", tab, cap, n, sink, sms);
		run<17, 1>("ld128 + returning add32 + CAS64", tab, cap, n, sink, sms);
		run<18, 1>("ld128 + CAS64 + red after", tab, cap, n, sink, sms);
		run<8, 1>("ld128 + red + CAS64", tab, cap, n, sink, sms);
		run<16, 2>("ld128 + CAS64 + returning add32", tab, cap, n, sink, sms);
	}
	return 0;
}
