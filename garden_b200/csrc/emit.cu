// Draw-list emission: sorted (key, payload) runs -> 64-byte records in the reference's UnsortedMesh / SortedMesh layout
// (include/garden/system/render/mesh.hpp:191-205), i.e. what renderUnsorted / renderSorted consume (mesh.cpp:556-770).
//   componentOffset = slot * componentStride       mesh.cpp:170,248
//   bakedModel      = (float4x3)model              mesh.cpp:171,249
//   distanceSq      = the key                      mesh.cpp:172,250-251
//   bufferIndex     = sorted-buffer index          mesh.cpp:252 (0 / padding for unsorted buffers)
#include "sceneprep_internal.h"
#include "sceneprep_math.cuh"
#include <algorithm>

namespace gsp
{

struct EmitArgs
{
	const SegmentDev* __restrict__ segments;
	const uint32_t* __restrict__ counters;
	const uint32_t* __restrict__ keys;
	uint32_t* __restrict__ payloads;      // in: pool << 28 | survivor index; out: pool << 28 | slot (what the exchange exports)
	gsp_record* __restrict__ records;
	const float4* world[kMaxPools];       // per survivor (compact)
	const uint32_t* surList[kMaxPools];   // survivor index -> slot
	uint32_t stride[kMaxPools];
	uint32_t bufferIndex[kMaxViews][kMaxPools];
	uint32_t segView[kMaxViews * kMaxPools];
	uint32_t segmentCount;
};

// Four lanes per record: lane q of a group produces the q-th 16-byte quarter, so a warp stores 8 records = 512 contiguous
// bytes per instruction; lanes 0..2 each gather one 16-byte third of the float4x3 world matrix and pass its upper half to
// the next lane by shuffle. Blocks walk their segment with a grid stride (segment lengths live on the device).
constexpr uint32_t kEmitThreads = 256, kEmitRecordsPerIter = kEmitThreads / 4, kEmitUnroll = 4;

__global__ void __launch_bounds__(kEmitThreads) kEmit(const __grid_constant__ EmitArgs A)
{
	// Record groups of ALL segments form one index space walked with a grid stride (list lengths differ by orders of
	// magnitude between views, and they are only known on the device).
	__shared__ uint32_t sGroupEnd[kMaxViews * kMaxPools];
	constexpr uint32_t kGroup = kEmitRecordsPerIter * kEmitUnroll;
	if (threadIdx.x == 0)
	{
		uint32_t running = 0;
		for (uint32_t i = 0; i < A.segmentCount; i++)
		{
			const SegmentDev sg = A.segments[i];
			const uint32_t c = sg.countIndex == kNone ? 0u : A.counters[sg.countIndex];
			running += (c + kGroup - 1) / kGroup;
			sGroupEnd[i] = running;
		}
	}
	__syncthreads();
	const uint32_t totalGroups = sGroupEnd[A.segmentCount - 1];
	const uint32_t q = threadIdx.x & 3;
	uint32_t segIndex = 0;
	for (uint32_t g = blockIdx.x; g < totalGroups; g += gridDim.x)
	{
		while (sGroupEnd[segIndex] <= g) segIndex++; // g only grows
		const SegmentDev seg = A.segments[segIndex];
		const uint32_t count = A.counters[seg.countIndex];
		const uint32_t view = A.segView[segIndex];
		const uint32_t* __restrict__ keys = A.keys + seg.offset;
		uint32_t* __restrict__ payloads = A.payloads + seg.offset;
		float4* __restrict__ out = reinterpret_cast<float4*>(A.records + seg.offset);
		const uint32_t base = (g - (segIndex ? sGroupEnd[segIndex - 1] : 0u)) * kGroup;
		// kEmitUnroll independent records per lane group: all list reads, then all gathers, then all stores, so that
		// several dependent-load chains are in flight per thread (the gather is latency-bound otherwise)
		uint32_t k[kEmitUnroll], payload[kEmitUnroll];
		#pragma unroll
		for (uint32_t r = 0; r < kEmitUnroll; r++)
		{
			const uint32_t j = base + r * kEmitRecordsPerIter + (threadIdx.x >> 2);
			k[r] = 0; payload[r] = 0;
			if (j < count)
			{
				k[r] = keys[j];
				payload[r] = payloads[j];
			}
		}
		float4 w[kEmitUnroll];
		uint32_t slotOf[kEmitUnroll]; // lane 3 of the group resolves the survivor index to the pool slot
		#pragma unroll
		for (uint32_t r = 0; r < kEmitUnroll; r++)
		{
			const uint32_t j = base + r * kEmitRecordsPerIter + (threadIdx.x >> 2);
			const uint32_t pool = payload[r] >> 28, index = payload[r] & 0x0FFFFFFFu;
			w[r] = make_float4(0.f, 0.f, 0.f, 0.f);
			slotOf[r] = 0;
			if (j < count)
			{
				if (q < 3)
					w[r] = A.world[pool][(size_t)index * kWorldStride + q];
				else
					slotOf[r] = A.surList[pool][index];
			}
		}
		#pragma unroll
		for (uint32_t r = 0; r < kEmitUnroll; r++)
		{
			const uint32_t j = base + r * kEmitRecordsPerIter + (threadIdx.x >> 2);
			const uint32_t pool = payload[r] >> 28;
			const uint32_t slot = __shfl_sync(0xffffffffu, slotOf[r], (threadIdx.x & 28u) | 3u);
			const float pz = __shfl_up_sync(0xffffffffu, w[r].z, 1);
			const float pw = __shfl_up_sync(0xffffffffu, w[r].w, 1);
			float4 o;
			if (q == 0)
			{
				const uint64_t componentOffset = (uint64_t)slot * A.stride[pool];
				o = make_float4(__uint_as_float((uint32_t)componentOffset), __uint_as_float((uint32_t)(componentOffset >> 32)),
					w[r].x, w[r].y);
			}
			else if (q < 3)
				o = make_float4(pz, pw, w[r].x, w[r].y);
			else
				o = make_float4(pz, pw, orderedToFloat(seg.descending ? ~k[r] : k[r]), __uint_as_float(A.bufferIndex[view][pool]));
			if (j < count)
			{
				out[(size_t)j * 4 + q] = o;
				if (q == 3)
					payloads[j] = (pool << 28) | slot; // the run now names slots (gsp_export_runs*, gsp_get_sorted_run_device)
			}
		}
	}
}

uint32_t launchEmit(Context& c)
{
	const uint32_t nseg = (uint32_t)c.segments.size();
	uint32_t maxCap = 0;
	for (auto& s : c.segments)
		maxCap = max(maxCap, s.capacity);
	if (nseg == 0 || maxCap == 0)
		return 0;
	EmitArgs A = {};
	A.segments = c.dSegments; A.counters = c.dCounters; A.keys = c.keys[0]; A.payloads = c.payloads[0];
	A.records = c.records; A.segmentCount = nseg;
	for (uint32_t p = 0; p < c.poolCount; p++)
	{
		A.world[p] = c.pools[p].world;
		A.surList[p] = c.pools[p].surList;
		A.stride[p] = c.pools[p].stride;
	}
	for (uint32_t v = 0; v < (uint32_t)c.views.size(); v++)
		for (uint32_t p = 0; p < c.poolCount; p++)
		{
			const uint32_t rt = c.pools[p].renderType; // SortedMesh::bufferIndex only exists in the shared lists (mesh.cpp:252)
			A.bufferIndex[v][p] = (rt == GSP_RT_TRANSLUCENT || rt == GSP_RT_UI) ? c.bufferIndexOf[v][p] : 0;
		}
	for (uint32_t s = 0; s < nseg; s++)
		A.segView[s] = c.segments[s].view;
	// persistent-style grid: enough blocks to fill the machine, each strides over its segment
	uint64_t capGroups = 0;
	for (auto& sgm : c.segments)
		capGroups += (sgm.capacity + kEmitRecordsPerIter * kEmitUnroll - 1) / (kEmitRecordsPerIter * kEmitUnroll);
	dim3 grid((uint32_t)std::min<uint64_t>(capGroups, c.smCount * 8u));
	kEmit<<<grid, kEmitThreads, 0, c.stream>>>(A);
	return 1;
}

} // namespace gsp
