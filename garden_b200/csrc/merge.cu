// Multi-GPU exchange step: k-way merge of the per-GPU sorted (key, payload) runs after the NCCL all-gather.
//
// The reference has no counterpart (one process, one std::sort per list, mesh.cpp:265-328); this is what makes the sharded
// path produce the same draw order as a single sort over all entities: every GPU culls + sorts its contiguous entity
// range, the runs are all-gathered over NVLink, and each GPU merges ONE key range of every list (splitters are taken from
// the longest run, which every rank holds after the gather, so all ranks agree without further communication).
// Position of an element = sum over the runs of (number of elements that precede it); ties are broken by (rank, payload),
// i.e. by global entity order, because shards are contiguous ranges in rank order and runs are already tie-ordered by
// payload (stable radix sort of a slot-ordered compaction).
#include "sceneprep_internal.h"
#include <algorithm>

namespace gsp
{

struct MergeArgs
{
	const uint32_t* __restrict__ keys;      // gathered: rank r's block starts at r * rankStride
	const uint32_t* __restrict__ payloads;
	const uint32_t* __restrict__ offsets;   // [ranks][lists] start of (rank, list) inside the rank's block
	const uint32_t* __restrict__ counts;    // [ranks][lists]
	uint32_t* __restrict__ bounds;          // [lists][ranks][2] sub-range [lo, hi) of each run that falls into my key range
	uint32_t* __restrict__ sliceInfo;       // [lists][2] global position of my slice start, slice length
	uint32_t* __restrict__ outKeys;         // my slice of every list, list l at outOffsets[l]
	uint32_t* __restrict__ outPayloads;
	uint8_t* __restrict__ outRanks;
	const uint32_t* __restrict__ outOffsets; // [lists]
	uint32_t rankStride, ranks, lists, myRank;
};

// first index in [0, n) with a[i] >= key (lower) or a[i] > key (upper)
__device__ __forceinline__ uint32_t lowerBound(const uint32_t* __restrict__ a, uint32_t n, uint32_t key)
{
	uint32_t lo = 0, hi = n;
	while (lo < hi)
	{
		const uint32_t mid = (lo + hi) >> 1;
		if (a[mid] < key) lo = mid + 1; else hi = mid;
	}
	return lo;
}
__device__ __forceinline__ uint32_t upperBound(const uint32_t* __restrict__ a, uint32_t n, uint32_t key)
{
	uint32_t lo = 0, hi = n;
	while (lo < hi)
	{
		const uint32_t mid = (lo + hi) >> 1;
		if (a[mid] <= key) lo = mid + 1; else hi = mid;
	}
	return lo;
}

// One thread per (list, run): the sub-range of the run inside my key range [splitLo, splitHi).
// Splitter k = key at position k * count / ranks of the longest run (k = 1..ranks-1); range 0 starts at -inf, the last
// ends at +inf. (If every run is empty there is nothing to split.)
__global__ void kMergeBounds(const __grid_constant__ MergeArgs A)
{
	const uint32_t list = blockIdx.x, run = threadIdx.x;
	if (run >= A.ranks)
		return;
	// splitter source: the longest run of this list (lowest rank on ties); every rank sees the same counts
	uint32_t src = 0;
	for (uint32_t r = 1; r < A.ranks; r++)
		if (A.counts[r * A.lists + list] > A.counts[src * A.lists + list]) src = r;
	const uint32_t count0 = A.counts[src * A.lists + list];
	const uint32_t* run0 = A.keys + (size_t)src * A.rankStride + A.offsets[src * A.lists + list];
	const uint32_t me = A.myRank;
	bool hasLo = me > 0 && count0 > 0, hasHi = me + 1 < A.ranks && count0 > 0;
	const uint32_t keyLo = hasLo ? run0[(uint64_t)me * count0 / A.ranks] : 0u;
	const uint32_t keyHi = hasHi ? run0[(uint64_t)(me + 1) * count0 / A.ranks] : 0u;
	const uint32_t n = A.counts[run * A.lists + list];
	const uint32_t* a = A.keys + (size_t)run * A.rankStride + A.offsets[run * A.lists + list];
	// a key equal to a splitter belongs to the range that starts at the splitter
	uint32_t lo = hasLo ? lowerBound(a, n, keyLo) : 0u;
	uint32_t hi = hasHi ? lowerBound(a, n, keyHi) : n;
	if (hi < lo) hi = lo; // equal splitters
	A.bounds[(list * A.ranks + run) * 2 + 0] = lo;
	A.bounds[(list * A.ranks + run) * 2 + 1] = hi;
	__syncthreads();
	if (run == 0)
	{
		uint32_t start = 0, length = 0;
		for (uint32_t r = 0; r < A.ranks; r++)
		{
			start += A.bounds[(list * A.ranks + r) * 2 + 0];
			length += A.bounds[(list * A.ranks + r) * 2 + 1] - A.bounds[(list * A.ranks + r) * 2 + 0];
		}
		A.sliceInfo[list * 2 + 0] = start;
		A.sliceInfo[list * 2 + 1] = length;
	}
}

// grid (blocks, ranks, lists): each thread places elements of one run's sub-range into my slice.
__global__ void __launch_bounds__(256) kMergeSlice(const __grid_constant__ MergeArgs A)
{
	const uint32_t list = blockIdx.z, run = blockIdx.y;
	const uint32_t lo = A.bounds[(list * A.ranks + run) * 2 + 0], hi = A.bounds[(list * A.ranks + run) * 2 + 1];
	const uint32_t sliceStart = A.sliceInfo[list * 2 + 0];
	const uint32_t* myKeys = A.keys + (size_t)run * A.rankStride + A.offsets[run * A.lists + list];
	const uint32_t* myPays = A.payloads + (size_t)run * A.rankStride + A.offsets[run * A.lists + list];
	const uint32_t outBase = A.outOffsets[list];
	for (uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x)
	{
		const uint32_t key = myKeys[i];
		uint32_t pos = i;
		for (uint32_t r = 0; r < A.ranks; r++)
		{
			if (r == run)
				continue;
			const uint32_t n = A.counts[r * A.lists + list];
			const uint32_t* a = A.keys + (size_t)r * A.rankStride + A.offsets[r * A.lists + list];
			// lower ranks win ties (they hold lower global entity indices)
			pos += r < run ? upperBound(a, n, key) : lowerBound(a, n, key);
		}
		const uint32_t o = outBase + (pos - sliceStart);
		A.outKeys[o] = key;
		A.outPayloads[o] = myPays[i];
		A.outRanks[o] = (uint8_t)run;
	}
}

uint32_t launchMerge(cudaStream_t stream, const MergeArgs& A, uint32_t maxRunLength)
{
	if (A.lists == 0 || A.ranks == 0)
		return 0;
	kMergeBounds<<<A.lists, 32, 0, stream>>>(A);
	const uint32_t blocks = std::max(1u, std::min((maxRunLength / A.ranks + 255u) / 256u + 1u, 148u * 4u / std::max(1u, A.ranks)));
	kMergeSlice<<<dim3(blocks, A.ranks, A.lists), 256, 0, stream>>>(A);
	return 2;
}

} // namespace gsp

using namespace gsp;

extern "C" int gsp_merge_gathered(void* cudaStream, uint32_t ranks, uint32_t myRank, uint32_t lists, uint32_t rankStride,
	const uint32_t* dKeys, const uint32_t* dPayloads, const uint32_t* dOffsets, const uint32_t* dCounts, uint32_t maxRunLength,
	uint32_t* dBounds, uint32_t* dSliceInfo, uint32_t* dOutKeys, uint32_t* dOutPayloads, uint8_t* dOutRanks,
	const uint32_t* dOutOffsets)
{
	if (ranks == 0 || ranks > 32 || myRank >= ranks || !dKeys || !dPayloads || !dOffsets || !dCounts || !dBounds ||
		!dSliceInfo || !dOutKeys || !dOutPayloads || !dOutRanks || !dOutOffsets)
		return GSP_ERR_INVALID;
	MergeArgs A;
	A.keys = dKeys; A.payloads = dPayloads; A.offsets = dOffsets; A.counts = dCounts; A.bounds = dBounds;
	A.sliceInfo = dSliceInfo; A.outKeys = dOutKeys; A.outPayloads = dOutPayloads; A.outRanks = dOutRanks;
	A.outOffsets = dOutOffsets; A.rankStride = rankStride; A.ranks = ranks; A.lists = lists; A.myRank = myRank;
	launchMerge((cudaStream_t)cudaStream, A, maxRunLength);
	return cudaGetLastError() == cudaSuccess ? GSP_OK : GSP_ERR_CUDA;
}
