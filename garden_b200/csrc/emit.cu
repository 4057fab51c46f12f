// Draw-list emission: sorted (key, payload) runs -> 64-byte records in the reference's UnsortedMesh / SortedMesh layout
// (include/garden/system/render/mesh.hpp:191-205), i.e. what renderUnsorted / renderSorted consume (mesh.cpp:556-770).
//   componentOffset = slot * componentStride       mesh.cpp:170,248
//   bakedModel      = (float4x3)model              mesh.cpp:171,249
//   distanceSq      = the key                      mesh.cpp:172,250-251
//   bufferIndex     = sorted-buffer index          mesh.cpp:252 (0 / padding for unsorted buffers)
#include "sceneprep_internal.h"
#include "sceneprep_math.cuh"

namespace gsp
{

struct EmitArgs
{
	const SegmentDev* __restrict__ segments;
	const uint32_t* __restrict__ counters;
	const uint32_t* __restrict__ keys;
	const uint32_t* __restrict__ payloads;
	gsp_record* __restrict__ records;
	const float4* world[kMaxPools];
	uint32_t stride[kMaxPools];
	uint32_t bufferIndex[kMaxViews][kMaxPools];
	uint32_t segView[kMaxViews * kMaxPools];
};

__global__ void __launch_bounds__(256) kEmit(const __grid_constant__ EmitArgs A)
{
	const SegmentDev seg = A.segments[blockIdx.y];
	const uint32_t count = seg.countIndex == kNone ? 0u : A.counters[seg.countIndex];
	const uint32_t j = blockIdx.x * 256 + threadIdx.x;
	if (j >= count)
		return;
	const uint32_t k = A.keys[seg.offset + j];
	const uint32_t payload = A.payloads[seg.offset + j];
	const uint32_t pool = payload >> 28, slot = payload & 0x0FFFFFFFu;
	const float4* w = A.world[pool] + (size_t)slot * 3;
	const float4 w0 = w[0], w1 = w[1], w2 = w[2];
	const float key = orderedToFloat(seg.descending ? ~k : k);
	const uint64_t componentOffset = (uint64_t)slot * A.stride[pool];
	const uint32_t bufferIndex = A.bufferIndex[A.segView[blockIdx.y]][pool];
	float4* out = reinterpret_cast<float4*>(A.records + seg.offset + j);
	out[0] = make_float4(__uint_as_float((uint32_t)componentOffset), __uint_as_float((uint32_t)(componentOffset >> 32)), w0.x, w0.y);
	out[1] = make_float4(w0.z, w0.w, w1.x, w1.y);
	out[2] = make_float4(w1.z, w1.w, w2.x, w2.y);
	out[3] = make_float4(w2.z, w2.w, key, __uint_as_float(bufferIndex));
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
	A.records = c.records;
	for (uint32_t p = 0; p < c.poolCount; p++)
	{
		A.world[p] = c.pools[p].world;
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
	dim3 grid((maxCap + 255) / 256, nseg);
	kEmit<<<grid, 256, 0, c.stream>>>(A);
	return 1;
}

} // namespace gsp
