// Segmented one-sweep LSD radix sort of (key, payload) pairs: 8-bit digits, 4 passes over 32-bit keys, stable.
//
// Replaces MeshRenderSystem::sortMeshes (source/system/render/mesh.cpp:265-328): one std::sort per unsorted buffer plus the
// translucent and UI lists. Here every list of every view is one segment and all segments are sorted by the same launches.
// Keys are order-preserving integer images of distanceSq (ascending lists) or their complement (descending lists,
// mesh.hpp:204), payload = pool << 28 | slot. The compaction emits each list in (pool, slot) order and the sort is stable,
// so equal keys keep ascending (pool, slot) order — the canonical tie-break (the reference's order on ties is unspecified).
//
// Algorithm (Adinets & Merrill, "Onesweep"): one up-front pass builds the 4 digit histograms of every segment; each pass
// then reads a tile once, ranks it in shared memory, resolves the tile's global digit offsets with a decoupled look-back
// over the preceding tiles and scatters. Per element traffic: 4 B (histogram) + 4 x (8 B read + 8 B write).
#include "sceneprep_internal.h"

namespace gsp
{

constexpr uint32_t kRadixBits = 8, kRadix = 1u << kRadixBits, kPasses = 4;
constexpr uint32_t kWarps = kSortThreads / 32;
constexpr uint32_t kFlagAggregate = 1u << 30, kFlagInclusive = 2u << 30, kValueMask = (1u << 30) - 1;

struct SortArgs
{
	const SegmentDev* __restrict__ segments;
	const uint32_t* __restrict__ counters; // segment length = counters[countIndex] (kNone -> 0)
	uint32_t* __restrict__ hist;           // [segment][pass][256]
	uint32_t* __restrict__ status;         // [pass][tileBase(segment) + tile][256]
	uint32_t* __restrict__ tickets;        // [segment][pass]
	const uint32_t* __restrict__ segTileOffset; // first status tile of each segment
	uint32_t tilesTotal;                   // status tiles per pass
};

__device__ __forceinline__ uint32_t segmentCount(const SortArgs& a, const SegmentDev& s)
{
	return s.countIndex == kNone ? 0u : a.counters[s.countIndex];
}
__device__ __forceinline__ uint32_t ldRelaxed(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void stRelease(uint32_t* p, uint32_t v)
{
	asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Up-front histogram of all 4 digits; also clears the look-back status words of the tile for all 4 passes.
__global__ void __launch_bounds__(kSortThreads) kSortHistogram(const __grid_constant__ SortArgs A,
	const uint32_t* __restrict__ keys)
{
	const SegmentDev seg = A.segments[blockIdx.y];
	const uint32_t count = segmentCount(A, seg);
	const uint32_t base = blockIdx.x * kSortTile;
	if (!seg.sorted || base >= count)
		return;
	__shared__ uint32_t sHist[kPasses][kRadix];
	for (uint32_t i = threadIdx.x; i < kPasses * kRadix; i += kSortThreads)
		(&sHist[0][0])[i] = 0;
	const uint32_t statusTile = A.segTileOffset[blockIdx.y] + blockIdx.x;
	#pragma unroll
	for (uint32_t p = 0; p < kPasses; p++)
		A.status[((size_t)p * A.tilesTotal + statusTile) * kRadix + threadIdx.x] = 0;
	__syncthreads();
	const uint32_t* k = keys + seg.offset + base;
	const uint32_t n = min(kSortTile, count - base);
	#pragma unroll 4
	for (uint32_t i = threadIdx.x; i < n; i += kSortThreads)
	{
		uint32_t key = k[i];
		atomicAdd(&sHist[0][key & 255u], 1u);
		atomicAdd(&sHist[1][(key >> 8) & 255u], 1u);
		atomicAdd(&sHist[2][(key >> 16) & 255u], 1u);
		atomicAdd(&sHist[3][key >> 24], 1u);
	}
	__syncthreads();
	uint32_t* h = A.hist + (size_t)blockIdx.y * kPasses * kRadix;
	for (uint32_t i = threadIdx.x; i < kPasses * kRadix; i += kSortThreads)
	{
		uint32_t c = (&sHist[0][0])[i];
		if (c)
			atomicAdd(&h[i], c);
	}
}

// Block-wide exclusive scan of one value per thread (kSortThreads == kRadix values).
__device__ __forceinline__ uint32_t blockExclusiveScan(uint32_t v, uint32_t* sWarpTotals /*[kWarps]*/)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= (uint32_t)o) inc += t;
	}
	if (lane == 31)
		sWarpTotals[warp] = inc;
	__syncthreads();
	uint32_t offset = 0;
	#pragma unroll
	for (uint32_t w = 0; w < kWarps; w++)
		if (w < warp) offset += sWarpTotals[w];
	__syncthreads();
	return offset + inc - v;
}

__global__ void __launch_bounds__(kSortThreads) kSortPass(const __grid_constant__ SortArgs A, uint32_t pass,
	const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ payIn,
	uint32_t* __restrict__ keysOut, uint32_t* __restrict__ payOut)
{
	const SegmentDev seg = A.segments[blockIdx.y];
	const uint32_t count = segmentCount(A, seg);
	if (!seg.sorted || blockIdx.x * kSortTile >= count)
		return;

	__shared__ uint32_t sKeys[kSortTile];
	__shared__ uint32_t sPay[kSortTile];
	__shared__ uint32_t sWarpHist[kWarps][kRadix];
	__shared__ uint32_t sBinStart[kRadix];  // tile-local exclusive start of each digit
	__shared__ int32_t sGlobal[kRadix];     // global position = sGlobal[d] + local index
	__shared__ uint32_t sWarpTotals[kWarps];
	__shared__ uint32_t sTile;

	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t shift = pass * kRadixBits;
	if (threadIdx.x == 0)
		sTile = atomicAdd(&A.tickets[blockIdx.y * kPasses + pass], 1u);
	for (uint32_t i = threadIdx.x; i < kWarps * kRadix; i += kSortThreads)
		(&sWarpHist[0][0])[i] = 0;
	__syncthreads();
	const uint32_t tile = sTile;
	const uint32_t base = tile * kSortTile;
	const uint32_t n = min(kSortTile, count - base);
	const uint32_t* kin = keysIn + seg.offset + base;
	const uint32_t* pin = payIn + seg.offset + base;

	// ---- load (warp-striped: tile order = warp, item, lane) and rank within the warp, in order ----
	uint32_t key[kSortItems], pay[kSortItems], rank[kSortItems];
	#pragma unroll
	for (uint32_t i = 0; i < kSortItems; i++)
	{
		uint32_t idx = warp * (kSortItems * 32) + i * 32 + lane;
		key[i] = idx < n ? kin[idx] : 0xFFFFFFFFu; // padding ranks after every real key of the tile
		pay[i] = idx < n ? pin[idx] : 0u;
	}
	#pragma unroll
	for (uint32_t i = 0; i < kSortItems; i++)
	{
		const uint32_t d = (key[i] >> shift) & (kRadix - 1);
		const uint32_t peers = __match_any_sync(0xffffffffu, d);
		const uint32_t leader = __ffs(peers) - 1;
		uint32_t pre = 0;
		if (lane == leader)
		{
			pre = sWarpHist[warp][d];
			sWarpHist[warp][d] = pre + __popc(peers);
		}
		pre = __shfl_sync(0xffffffffu, pre, leader);
		rank[i] = pre + __popc(peers & ((1u << lane) - 1u));
		__syncwarp();
	}
	__syncthreads();

	// ---- per digit (thread d owns digit d): warp offsets, tile count, global offset by look-back ----
	const uint32_t d = threadIdx.x;
	uint32_t tileCount = 0;
	#pragma unroll
	for (uint32_t w = 0; w < kWarps; w++)
	{
		uint32_t c = sWarpHist[w][d];
		sWarpHist[w][d] = tileCount;
		tileCount += c;
	}
	const uint32_t binStart = blockExclusiveScan(tileCount, sWarpTotals);
	// global exclusive start of digit d in this pass = exclusive scan of the segment histogram
	const uint32_t histD = A.hist[((size_t)blockIdx.y * kPasses + pass) * kRadix + d];
	const uint32_t digitBase = blockExclusiveScan(histD, sWarpTotals);

	// padding keys (digit 255 in every pass) are excluded from what is published to other tiles
	uint32_t realCount = tileCount;
	if (d == kRadix - 1)
		realCount -= kSortTile - n;
	uint32_t* st = A.status + ((size_t)pass * A.tilesTotal + A.segTileOffset[blockIdx.y] + tile) * kRadix + d;
	uint32_t exclusive = 0;
	if (tile == 0)
		stRelease(st, kFlagInclusive | realCount);
	else
	{
		stRelease(st, kFlagAggregate | realCount);
		int32_t t = (int32_t)tile - 1;
		while (true)
		{
			const uint32_t* ps = st - (size_t)(tile - (uint32_t)t) * kRadix;
			uint32_t s;
			do { s = ldRelaxed(ps); } while ((s & ~kValueMask) == 0);
			exclusive += s & kValueMask;
			if (s & kFlagInclusive)
				break;
			t--;
		}
		stRelease(st, kFlagInclusive | (exclusive + realCount));
	}
	sBinStart[d] = binStart;
	sGlobal[d] = (int32_t)(digitBase + exclusive) - (int32_t)binStart;
	__syncthreads();

	// ---- scatter into shared memory in digit order, then write runs out ----
	#pragma unroll
	for (uint32_t i = 0; i < kSortItems; i++)
	{
		const uint32_t dg = (key[i] >> shift) & (kRadix - 1);
		const uint32_t local = sBinStart[dg] + sWarpHist[warp][dg] + rank[i];
		sKeys[local] = key[i];
		sPay[local] = pay[i];
	}
	__syncthreads();
	uint32_t* kout = keysOut + seg.offset;
	uint32_t* pout = payOut + seg.offset;
	#pragma unroll 4
	for (uint32_t j = threadIdx.x; j < n; j += kSortThreads)
	{
		const uint32_t k = sKeys[j];
		const uint32_t dg = (k >> shift) & (kRadix - 1);
		const uint32_t pos = (uint32_t)(sGlobal[dg] + (int32_t)j);
		kout[pos] = k;
		pout[pos] = sPay[j];
	}
}

uint32_t launchSort(Context& c, cudaEvent_t afterHistogram)
{
	const uint32_t nseg = (uint32_t)c.segments.size();
	if (nseg == 0)
	{
		if (afterHistogram) cudaEventRecord(afterHistogram, c.stream);
		return 0;
	}
	uint32_t maxTiles = 0;
	bool anySorted = false;
	for (auto& s : c.segments)
	{
		if (!s.sorted) continue;
		anySorted = true;
		maxTiles = max(maxTiles, (s.capacity + kSortTile - 1) / kSortTile);
	}
	if (!anySorted || maxTiles == 0)
	{
		if (afterHistogram) cudaEventRecord(afterHistogram, c.stream);
		return 0;
	}
	SortArgs A;
	A.segments = c.dSegments; A.counters = c.dCounters; A.hist = c.sortHist; A.status = c.sortStatus;
	A.tickets = c.sortTickets; A.segTileOffset = c.segTileOffset; A.tilesTotal = c.sortTilesTotal;
	cudaMemsetAsync(c.sortHist, 0, (size_t)nseg * kPasses * kRadix * sizeof(uint32_t), c.stream);
	cudaMemsetAsync(c.sortTickets, 0, (size_t)nseg * kPasses * sizeof(uint32_t), c.stream);
	dim3 grid(maxTiles, nseg);
	kSortHistogram<<<grid, kSortThreads, 0, c.stream>>>(A, c.keys[0]);
	if (afterHistogram) cudaEventRecord(afterHistogram, c.stream);
	uint32_t launches = 1;
	for (uint32_t pass = 0; pass < kPasses; pass++)
	{
		const uint32_t in = pass & 1, out = in ^ 1;
		kSortPass<<<grid, kSortThreads, 0, c.stream>>>(A, pass, c.keys[in], c.payloads[in], c.keys[out], c.payloads[out]);
		launches++;
	}
	return launches; // 4 passes: sorted data ends in buffer 0
}

} // namespace gsp
