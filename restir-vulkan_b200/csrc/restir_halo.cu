// restir_halo.cu — halo exchange of the row-band split as this library's own kernels over NVLink peer memory.
//
// A band context whose neighbours are connected (restir_band_connect) pushes the reservoir rows a neighbour holds
// as halo straight into that neighbour's buffer when the pass that produced them has finished — plain stores to
// peer memory (cudaIpc mapping in the multi-process case, raw pointers between contexts of one process) — and then
// raises a per-(side, buffer) counter in the neighbour's flag block.  The pass that is about to read a halo first
// waits, on the device, until both neighbours' counters have reached the number of times this buffer has been
// produced (every rank runs the same pass sequence).  No host synchronisation, no collective, no copy engine.
#include "restir_kernels.h"

namespace restir {

// Copies `rows` boundary rows to each connected neighbour and signals them.  One launch per produced buffer.
__global__ void __launch_bounds__(256) halo_push_kernel(HaloPush hp) {
	const size_t rowVec = (size_t)hp.W * (sizeof(PackedReservoir) / sizeof(uint4)); // uint4 per row
	const uint4 *src = reinterpret_cast<const uint4 *>(hp.local);
	for (int side = 0; side < 2; ++side) {
		uint4 *dst = reinterpret_cast<uint4 *>(hp.peer[side]);
		if (dst == nullptr || hp.rows[side] <= 0) {
			continue;
		}
		// rows [first, first + rows) of the screen: local and peer buffers start at their own allocBegin
		const size_t n = (size_t)hp.rows[side] * rowVec;
		const size_t so = (size_t)(hp.firstRow[side] - hp.localAllocBegin) * rowVec, to = (size_t)(hp.firstRow[side] - hp.peerAllocBegin[side]) * rowVec;
		for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
			dst[to + i] = src[so + i];
		}
	}
	// the last block to finish raises the neighbours' counters: every store above is visible system-wide before it
	__threadfence_system();
	__shared__ unsigned last;
	__syncthreads();
	if (threadIdx.x == 0) {
		last = atomicAdd(hp.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
	}
	__syncthreads();
	if (last && threadIdx.x == 0) {
		*hp.ticket = 0u;
		__threadfence_system();
		for (int side = 0; side < 2; ++side) {
			if (hp.peerFlag[side] != nullptr) {
				*reinterpret_cast<volatile unsigned long long *>(hp.peerFlag[side]) = hp.sequence;
			}
		}
		__threadfence_system();
	}
}

// Spins until both flags have reached `sequence` (or a 5 s timeout, counted: a lost neighbour must not hang the GPU).
__global__ void halo_wait_kernel(const unsigned long long *flagA, const unsigned long long *flagB, unsigned long long sequence,
                                 unsigned long long timeoutNs, unsigned long long *counters) {
	unsigned long long start;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(start));
	for (;;) {
		bool a = flagA == nullptr || *reinterpret_cast<const volatile unsigned long long *>(flagA) >= sequence;
		bool b = flagB == nullptr || *reinterpret_cast<const volatile unsigned long long *>(flagB) >= sequence;
		if (a && b) {
			break;
		}
		unsigned long long now;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
		if (now - start > timeoutNs) {
			atomicAdd(counters + kCounterHaloTimeout, 1ull);
			break;
		}
		__nanosleep(200);
	}
	__threadfence_system();
}

cudaError_t launch_halo_push(const HaloPush &hp, int smCount, cudaStream_t s) {
	size_t vecs = (size_t)(hp.rows[0] > hp.rows[1] ? hp.rows[0] : hp.rows[1]) * hp.W * 2;
	unsigned grid = (unsigned)((vecs + 256 * 8 - 1) / (256 * 8));
	grid = grid < 1 ? 1 : (grid > (unsigned)smCount * 2 ? (unsigned)smCount * 2 : grid);
	halo_push_kernel<<<grid, 256, 0, s>>>(hp);
	return cudaGetLastError();
}
cudaError_t launch_halo_wait(const unsigned long long *flagA, const unsigned long long *flagB, unsigned long long sequence, unsigned long long *counters,
                             cudaStream_t s) {
	halo_wait_kernel<<<1, 1, 0, s>>>(flagA, flagB, sequence, 5000000000ull, counters);
	return cudaGetLastError();
}

cudaError_t preload_halo_kernels() {
	cudaFuncAttributes a;
	cudaError_t e = cudaFuncGetAttributes(&a, halo_push_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, halo_wait_kernel);
	return e;
}

} // namespace restir
