// isVisible write-back: MeshRenderComponent::isVisible (offset 15) is the one field the path stores INTO the ECS
// (source/system/render/mesh.cpp:144-146,152-153,161-167). The host pool is AoS, so the bytes land `stride` apart:
//   full  — the per-slot bytes are packed to one bit per slot on the device (32x less PCIe traffic than the byte array) and
//           scattered into the pool by a few host threads;
//   delta — the device knows which byte the host holds for every slot (uploaded with the pool by kStagePool, or written by
//           the previous write-back), so only the slots whose value CHANGED travel: a compacted list of slot | new << 31.
// Both kernels store their (small) output straight into mapped pinned host memory: no copy-engine transfer is involved, so
// the write-back does not queue behind the draw lists that are travelling device -> host at the same time.
#include "sceneprep_internal.h"

namespace gsp
{

// One thread per slot: ballot -> one bit per slot. Also records that the host will hold exactly these bytes.
__global__ void __launch_bounds__(256) kPackVisible(uint32_t occupancy, const uint8_t* __restrict__ visible,
	uint8_t* __restrict__ flags, uint32_t* __restrict__ bits)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool in = i < occupancy;
	const bool v = in && visible[i] != 0;
	const uint32_t word = __ballot_sync(0xffffffffu, v);
	if ((threadIdx.x & 31) == 0 && in)
		bits[i >> 5] = word;
	if (in)
	{
		const uint8_t f = flags[i];
		const uint8_t g = (uint8_t)((f & ~(kMfHostVisible | kMfHostVisibleOdd)) | (v ? kMfHostVisible : 0));
		if (g != f)
			flags[i] = g;
	}
}

// One thread per slot: slots whose new value differs from the byte the host holds are appended (warp-aggregated) to `list`.
__global__ void __launch_bounds__(256) kVisibleDelta(uint32_t occupancy, const uint8_t* __restrict__ visible,
	uint8_t* __restrict__ flags, uint32_t* __restrict__ list, uint32_t* __restrict__ count)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool in = i < occupancy;
	uint8_t f = 0;
	bool v = false, changed = false;
	if (in)
	{
		f = flags[i];
		v = visible[i] != 0;
		changed = (f & kMfHostVisibleOdd) || (((f & kMfHostVisible) != 0) != v);
	}
	const uint32_t mask = __ballot_sync(0xffffffffu, changed);
	if (mask == 0)
		return;
	const uint32_t lane = threadIdx.x & 31;
	uint32_t base = 0;
	if (lane == 0)
		base = atomicAdd(count, __popc(mask));
	base = __shfl_sync(0xffffffffu, base, 0);
	if (changed)
	{
		list[base + __popc(mask & ((1u << lane) - 1u))] = i | (v ? 0x80000000u : 0u);
		flags[i] = (uint8_t)((f & ~(kMfHostVisible | kMfHostVisibleOdd)) | (v ? kMfHostVisible : 0));
	}
}

uint32_t launchPackVisible(Context& c, uint32_t pool, uint32_t* dBits)
{
	auto& p = c.pools[pool];
	kPackVisible<<<(p.occupancy + 255) / 256, 256, 0, c.stream>>>(p.occupancy, p.visible, p.flags, dBits);
	return 1;
}

__global__ void kPublishWord(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst)
{
	*dst = *src;
}

uint32_t launchVisibleDelta(Context& c, uint32_t pool, uint32_t* list, uint32_t* dCount, uint32_t* hCountMapped)
{
	auto& p = c.pools[pool];
	kVisibleDelta<<<(p.occupancy + 255) / 256, 256, 0, c.stream>>>(p.occupancy, p.visible, p.flags, list, dCount);
	kPublishWord<<<1, 1, 0, c.stream>>>(dCount, hCountMapped);
	return 2;
}

} // namespace gsp
