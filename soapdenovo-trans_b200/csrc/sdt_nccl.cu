// sdt_nccl.cu — the super-k-mer exchange of the sliced build behind the C ABI: one call per epoch does what a
// host would otherwise script around sdtgpu_skm_stage / sdtgpu_skm_import (include/sdtgpu.h): the ordinal
// bound (an all-reduce), the merge and pack by owner, the counts (an all-gather), the records (ONE grouped
// ncclSend / ncclRecv straight into the import buffer) and the build of this rank's slices.
//
// Host code only, layered on the public C ABI of the handle.  NCCL is bound at run time (dlopen of
// libnccl.so.2: in a process that already holds an NCCL — torch's — that one is used; a plain C host gets the
// system's), so libsdtgpu.so has no link-time dependency on it and single-GPU users never load it.
// The reference has no counterpart: its workers share one address space (prlHashReads.c:79-88).
#include "../../include/sdtgpu.h"
#include <cuda_runtime.h>
#include <nccl.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Nccl
{
	void *lib = nullptr;
	ncclResult_t (*GetUniqueId) (ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank) (ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy) (ncclComm_t) = nullptr;
	const char *(*GetErrorString) (ncclResult_t) = nullptr;
	ncclResult_t (*AllReduce) (const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather) (const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send) (const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv) (void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart) () = nullptr;
	ncclResult_t (*GroupEnd) () = nullptr;
	std::string err;
};

Nccl g_nccl;
std::string g_comm_err;	// last failure without a communicator

template <class F> bool sym (void *lib, const char *name, F &f)
{
	f = reinterpret_cast<F> (dlsym (lib, name));
	return f != nullptr;
}

bool nccl_load ()
{
	Nccl &n = g_nccl;
	if (n.lib)
		return true;
	const char *names[] = { "libnccl.so.2", "libnccl.so" };
	void *lib = nullptr;
	for (const char *nm : names)
		if ((lib = dlopen (nm, RTLD_NOW | RTLD_LOCAL)))
			break;
	if (!lib)
	{
		const char *why = dlerror ();
		n.err = std::string ("cannot load libnccl.so.2: ") + (why ? why : "?");
		return false;
	}
	const bool ok = sym (lib, "ncclGetUniqueId", n.GetUniqueId) && sym (lib, "ncclCommInitRank", n.CommInitRank) && sym (lib, "ncclCommDestroy", n.CommDestroy)
		&& sym (lib, "ncclGetErrorString", n.GetErrorString) && sym (lib, "ncclAllReduce", n.AllReduce) && sym (lib, "ncclAllGather", n.AllGather)
		&& sym (lib, "ncclSend", n.Send) && sym (lib, "ncclRecv", n.Recv) && sym (lib, "ncclGroupStart", n.GroupStart) && sym (lib, "ncclGroupEnd", n.GroupEnd);
	if (!ok)
	{
		n.err = "libnccl.so.2 lacks a symbol of the point-to-point API (NCCL >= 2.7 is needed)";
		dlclose (lib);
		return false;
	}
	n.lib = lib;
	return true;
}

}	// namespace

struct sdtgpu_comm
{
	ncclComm_t comm = nullptr;
	int device = 0, rank = 0, world = 1;
	unsigned long long *d_buf = nullptr, *h_pin = nullptr;	// [0] mine, [1] reduced, [8 ..) my counts, [8 + world ..) everybody's
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	std::string err;
};

namespace {

int comm_fail (sdtgpu_comm *c, int code, const std::string &what)
{
	(c ? c->err : g_comm_err) = what;
	return code;
}

#define NC(c, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return comm_fail (c, SDTGPU_ECUDA, std::string (#call) + ": " + g_nccl.GetErrorString (r_)); } while (0)
#define CU(c, call) do { cudaError_t r_ = (call); if (r_ != cudaSuccess) return comm_fail (c, SDTGPU_ECUDA, std::string (#call) + ": " + cudaGetErrorString (r_)); } while (0)

}	// namespace

extern "C" {

const char *sdtgpu_comm_last_error (const sdtgpu_comm_t *c)
{
	return c ? c->err.c_str () : g_comm_err.c_str ();
}

int sdtgpu_comm_unique_id (uint8_t id[SDTGPU_COMM_ID_BYTES])
{
	if (!id)
		return SDTGPU_EINVAL;
	if (!nccl_load ())
		return comm_fail (nullptr, SDTGPU_ESTATE, g_nccl.err);
	static_assert (sizeof (ncclUniqueId) == SDTGPU_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
	ncclUniqueId u;
	NC (nullptr, g_nccl.GetUniqueId (&u));
	memcpy (id, &u, sizeof u);
	return SDTGPU_OK;
}

int sdtgpu_comm_create (sdtgpu_comm_t **out, int device, const uint8_t id[SDTGPU_COMM_ID_BYTES], int rank, int world)
{
	if (!out || !id || world < 1 || world > 64 || rank < 0 || rank >= world)
		return comm_fail (nullptr, SDTGPU_EINVAL, "sdtgpu_comm_create: need 0 <= rank < world <= 64");
	*out = nullptr;
	if (!nccl_load ())
		return comm_fail (nullptr, SDTGPU_ESTATE, g_nccl.err);
	sdtgpu_comm *c = new sdtgpu_comm;
	c->device = device;
	c->rank = rank;
	c->world = world;
	auto bail = [&](int rc) { g_comm_err = c->err; sdtgpu_comm_destroy (c); return rc; };
	if (cudaSetDevice (device) != cudaSuccess)
		return bail (comm_fail (c, SDTGPU_ECUDA, "cudaSetDevice failed"));
	ncclUniqueId u;
	memcpy (&u, id, sizeof u);
	ncclResult_t r = g_nccl.CommInitRank (&c->comm, world, u, rank);
	if (r != ncclSuccess)
	{
		c->comm = nullptr;
		return bail (comm_fail (c, SDTGPU_ECUDA, std::string ("ncclCommInitRank: ") + g_nccl.GetErrorString (r)));
	}
	const size_t words = 8 + (size_t) world + (size_t) world * world;
	if (cudaMalloc (&c->d_buf, words * 8) != cudaSuccess || cudaMallocHost (&c->h_pin, words * 8) != cudaSuccess
	    || cudaEventCreate (&c->e0) != cudaSuccess || cudaEventCreate (&c->e1) != cudaSuccess)
		return bail (comm_fail (c, SDTGPU_ENOMEM, "sdtgpu_comm_create: allocation failed"));
	*out = c;
	return SDTGPU_OK;
}

int sdtgpu_comm_destroy (sdtgpu_comm_t *c)
{
	if (!c)
		return SDTGPU_OK;
	cudaSetDevice (c->device);
	if (c->comm && g_nccl.CommDestroy)
		g_nccl.CommDestroy (c->comm);
	if (c->d_buf)
		cudaFree (c->d_buf);
	if (c->h_pin)
		cudaFreeHost (c->h_pin);
	if (c->e0)
		cudaEventDestroy (c->e0);
	if (c->e1)
		cudaEventDestroy (c->e1);
	delete c;
	return SDTGPU_OK;
}

int sdtgpu_skm_exchange (sdtgpu_t *h, sdtgpu_comm_t *c, uint64_t reads_end_this_rank, uint64_t *n_received, double *collective_ms)
{
	if (!h || !c)
		return SDTGPU_EINVAL;
	const int world = c->world, rank = c->rank;
	cudaStream_t s = static_cast<cudaStream_t> (sdtgpu_stream (h));
	CU (c, cudaSetDevice (c->device));
	int rc;
	// 1. reads of all ranks this epoch: 32-bit ordinals in the slice images when they fit
	c->h_pin[0] = reads_end_this_rank;
	CU (c, cudaMemcpyAsync (c->d_buf, c->h_pin, 8, cudaMemcpyHostToDevice, s));
	NC (c, g_nccl.AllReduce (c->d_buf, c->d_buf + 1, 1, ncclUint64, ncclMax, c->comm, s));
	CU (c, cudaMemcpyAsync (c->h_pin + 1, c->d_buf + 1, 8, cudaMemcpyDeviceToHost, s));
	CU (c, cudaStreamSynchronize (s));
	if ((rc = sdtgpu_skm_set_ordinal_bound (h, c->h_pin[1])))
		return comm_fail (c, rc, std::string ("sdtgpu_skm_set_ordinal_bound: ") + sdtgpu_last_error (h));
	// 2. this rank's copies merged, survivors packed by owner
	void *d_rec = nullptr;
	std::vector<uint64_t> starts (world), counts (world), recv (world);
	if ((rc = sdtgpu_skm_stage (h, &d_rec, starts.data (), counts.data ())))
		return comm_fail (c, rc, std::string ("sdtgpu_skm_stage: ") + sdtgpu_last_error (h));
	uint64_t geo[12];
	if ((rc = sdtgpu_slice_geometry (h, geo)))
		return comm_fail (c, rc, "sdtgpu_slice_geometry failed");
	const size_t rec_bytes = (size_t) geo[4], words = rec_bytes / 8;
	// 3. who sends how much to whom
	unsigned long long *h_mine = c->h_pin + 8, *h_all = c->h_pin + 8 + world, *d_mine = c->d_buf + 8, *d_all = c->d_buf + 8 + world;
	for (int r = 0; r < world; r++)
		h_mine[r] = counts[r];
	CU (c, cudaEventRecord (c->e0, s));
	CU (c, cudaMemcpyAsync (d_mine, h_mine, (size_t) world * 8, cudaMemcpyHostToDevice, s));
	NC (c, g_nccl.AllGather (d_mine, d_all, (size_t) world, ncclUint64, c->comm, s));
	CU (c, cudaMemcpyAsync (h_all, d_all, (size_t) world * world * 8, cudaMemcpyDeviceToHost, s));
	CU (c, cudaStreamSynchronize (s));
	uint64_t total = 0;
	for (int src = 0; src < world; src++)
		total += (recv[src] = h_all[(size_t) src * world + rank]);
	// 4. the records, in place: sender's region -> receiver's import buffer (runs arrive in source-rank order)
	void *d_in = nullptr;
	if ((rc = sdtgpu_skm_import_buffer (h, total, &d_in)))
		return comm_fail (c, rc, std::string ("sdtgpu_skm_import_buffer: ") + sdtgpu_last_error (h));
	const char *out_base = static_cast<const char *> (d_rec);
	char *in_base = static_cast<char *> (d_in);
	NC (c, g_nccl.GroupStart ());
	uint64_t off = 0;
	for (int src = 0; src < world; src++)
	{
		const uint64_t n = recv[src];
		char *seg = in_base + off * rec_bytes;
		off += n;
		if (!n)
			continue;
		if (src == rank)	// own records: a device-local copy, never touches the fabric
			CU (c, cudaMemcpyAsync (seg, out_base + starts[rank] * rec_bytes, n * rec_bytes, cudaMemcpyDeviceToDevice, s));
		else
			NC (c, g_nccl.Recv (seg, n * words, ncclUint64, src, c->comm, s));
	}
	for (int dst = 0; dst < world; dst++)
		if (dst != rank && counts[dst])
			NC (c, g_nccl.Send (out_base + starts[dst] * rec_bytes, counts[dst] * words, ncclUint64, dst, c->comm, s));
	NC (c, g_nccl.GroupEnd ());
	CU (c, cudaEventRecord (c->e1, s));
	// 5. this rank's slices
	if ((rc = sdtgpu_skm_import (h, total)))
		return comm_fail (c, rc, std::string ("sdtgpu_skm_import: ") + sdtgpu_last_error (h));
	if (n_received)
		*n_received = total;
	if (collective_ms)
	{
		float ms = 0;
		CU (c, cudaEventSynchronize (c->e1));
		CU (c, cudaEventElapsedTime (&ms, c->e0, c->e1));
		*collective_ms = ms;
	}
	return SDTGPU_OK;
}

}	// extern "C"
