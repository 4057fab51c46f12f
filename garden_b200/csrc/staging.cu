// Component staging: AoS ECS pools (as uploaded raw from the host) -> SoA mirrors in HBM.
//
// Replaces, once per structural change instead of per entity per view, the reference's
//   Manager::tryGet<TransformComponent>(entity)  (libraries/ecsm/include/ecsm.hpp:898-905, Entity::findComponent :274-278)
// with a device-resident entity -> transform-slot map, and the AoS field reads of
//   TransformComponent  (include/garden/system/transform.hpp:31-60)   and
//   MeshRenderComponent (include/garden/system/render/mesh.hpp:45-55).
#include "sceneprep_internal.h"
#include "sceneprep_math.cuh"

namespace gsp
{

__device__ __forceinline__ uint32_t ldU32(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ float ldF32(const uint8_t* p) { return *reinterpret_cast<const float*>(p); }

__global__ void __launch_bounds__(256) kMaxEntity(const uint8_t* __restrict__ aos, uint32_t stride, uint32_t count,
	uint32_t* __restrict__ maxOut)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t e = i < count ? ldU32(aos + (size_t)i * stride + kTfEntity) : 0;
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
	if ((threadIdx.x & 31) == 0 && e)
		atomicMax(maxOut, e);
}

// One thread per transform slot. `full` also (re)builds the hierarchy inputs and the entity map.
__global__ void __launch_bounds__(256) kStageTransforms(const uint8_t* __restrict__ aos, uint32_t stride,
	uint32_t first, uint32_t count, int full, float4* __restrict__ rot, float4* __restrict__ posSx,
	float2* __restrict__ sYZ, uint16_t* __restrict__ flags, uint32_t* __restrict__ entity,
	uint32_t* __restrict__ parentEntity, uint32_t* __restrict__ entityToSlot)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	uint32_t slot = first + i;
	const uint8_t* t = aos + (size_t)i * stride;
	uint32_t e = ldU32(t + kTfEntity);
	const float4 q = make_float4(ldF32(t + kTfRot), ldF32(t + kTfRot + 4), ldF32(t + kTfRot + 8), ldF32(t + kTfRot + 12));
	const float4 ps = make_float4(ldF32(t + kTfPos), ldF32(t + kTfPos + 4), ldF32(t + kTfPos + 8), ldF32(t + kTfScale));
	const float2 syz = make_float2(ldF32(t + kTfScale + 4), ldF32(t + kTfScale + 8));
	rot[slot] = q; posSx[slot] = ps; sYZ[slot] = syz;
	uint32_t w = ldU32(t + kTfSelfActive); // bytes 72..75: selfActive, ancestorsActive, modelWithAncestors, pad
	uint16_t f = 0;
	Mat43 unused;
	if (!localModel43Fast<true>(ps.x, ps.y, ps.z, q.x, q.y, q.z, q.w, ps.w, syz.x, syz.y, unused))
		f |= kTfExactLocal; // the per-frame kernel must use the guarded 4-lane code for this transform
	if (e) f |= kTfLive;
	if ((w & 0xffu) && (w & 0xff00u)) f |= kTfActive;   // isActive(), transform.hpp:110
	if (w & 0xff0000u) f |= kTfAncestors;                // modelWithAncestors, transform.hpp:60,200
	// the upper byte holds the chain length (kComputeDepth); a TRS-only update keeps it
	flags[slot] = full ? f : (uint16_t)(f | (flags[slot] & ~((1u << kTfDepthShift) - 1u)));
	if (full)
	{
		entity[slot] = e;
		parentEntity[slot] = e ? ldU32(t + kTfParent) : 0;
		if (e)
			entityToSlot[e] = slot + 1;
	}
}

// parent entity id -> parent transform slot. A live parent id without a TransformComponent is an error
// (the reference's Manager::get throws, ecsm.hpp:863-873).
__global__ void __launch_bounds__(256) kResolveParents(uint32_t count, const uint32_t* __restrict__ parentEntity,
	const uint32_t* __restrict__ entityToSlot, uint32_t entityCap, uint32_t* __restrict__ parent, uint32_t* __restrict__ error)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	uint32_t pe = parentEntity[i];
	uint32_t p = kNone;
	if (pe)
	{
		uint32_t s = pe < entityCap ? entityToSlot[pe] : 0;
		if (s) p = s - 1;
		else atomicExch(error, (uint32_t)GSP_ERR_HIERARCHY);
	}
	parent[i] = p;
}

// Chain length of every transform (number of ancestors, saturating at 255) into the upper flag byte. The cull kernel uses
// it as the trip count of its chain loop and to group work of equal chain length into the same warp; a wrong value only
// costs time (the links decide), never correctness.
__global__ void __launch_bounds__(256) kComputeDepth(uint32_t count, const uint32_t* __restrict__ parent, uint16_t* __restrict__ flags)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	uint32_t depth = 0;
	uint32_t p = parent[i];
	while (p != kNone && depth < kTfDepthMax)
	{
		depth++;
		p = parent[p];
	}
	flags[i] = (uint16_t)((flags[i] & ((1u << kTfDepthShift) - 1u)) | (depth << kTfDepthShift));
}

// One thread per mesh-component slot: AABB, owner entity and the static part of the filter at mesh.cpp:140-147.
__global__ void __launch_bounds__(256) kStagePool(const uint8_t* __restrict__ aos, uint32_t stride, uint32_t count,
	float4* __restrict__ aabbA, float2* __restrict__ aabbB, uint32_t* __restrict__ entity, uint8_t* __restrict__ flags)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const uint8_t* m = aos + (size_t)i * stride;
	uint32_t e = ldU32(m + kMcEntity);
	uint32_t w = ldU32(m + 12); // bytes 12..15: reserved2 (u16), isEnabled, isVisible
	float mnx = ldF32(m + kMcAabbMin), mny = ldF32(m + kMcAabbMin + 4), mnz = ldF32(m + kMcAabbMin + 8);
	float mxx = ldF32(m + kMcAabbMax), mxy = ldF32(m + kMcAabbMax + 4), mxz = ldF32(m + kMcAabbMax + 8);
	aabbA[i] = make_float4(mnx, mny, mnz, mxx);
	aabbB[i] = make_float2(mxy, mxz);
	entity[i] = e;
	// aabb.getSize() = max - min, fixW(), areAllTrue(size <= 0)  (mesh.cpp:140-142, aabb.hpp:142)
	bool degenerate = (__fsub_rn(mxx, mnx) <= 0.0f) && (__fsub_rn(mxy, mny) <= 0.0f) && (__fsub_rn(mxz, mnz) <= 0.0f);
	bool enabled = (w & 0xff0000u) != 0;
	flags[i] = (e && enabled && !degenerate) ? kMfCandidate : 0;
}

__global__ void __launch_bounds__(256) kLinkPool(uint32_t count, const uint32_t* __restrict__ entity,
	const uint32_t* __restrict__ entityToSlot, uint32_t entityCap, uint32_t* __restrict__ tslot)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	uint32_t e = entity[i];
	uint32_t s = (e && e < entityCap) ? entityToSlot[e] : 0;
	tslot[i] = s ? s - 1 : kNone;
}

static inline uint32_t blocksFor(uint32_t n, uint32_t per) { return (n + per - 1) / per; }

uint32_t launchStageTransforms(Context& c, const void* dAos, uint32_t stride, uint32_t first, uint32_t count, bool full)
{
	if (count == 0)
		return 0;
	auto& t = c.tf;
	kStageTransforms<<<blocksFor(count, 256), 256, 0, c.stream>>>((const uint8_t*)dAos, stride, first, count, full ? 1 : 0,
		t.rot, t.posSx, t.sYZ, t.flags, t.entity, t.parentEntity, t.entityToSlot);
	uint32_t n = 1;
	if (full)
	{
		kResolveParents<<<blocksFor(count, 256), 256, 0, c.stream>>>(count, t.parentEntity, t.entityToSlot, t.entityCap,
			t.parent, c.dError);
		kComputeDepth<<<blocksFor(count, 256), 256, 0, c.stream>>>(count, t.parent, t.flags);
		n += 2;
	}
	return n;
}

uint32_t launchMaxEntity(Context& c, const void* dAos, uint32_t stride, uint32_t count, uint32_t* dMax)
{
	if (count == 0)
		return 0;
	kMaxEntity<<<blocksFor(count, 256), 256, 0, c.stream>>>((const uint8_t*)dAos, stride, count, dMax);
	return 1;
}

uint32_t launchStagePool(Context& c, uint32_t pool, const void* dAos, uint32_t stride, uint32_t occupancy)
{
	if (occupancy == 0)
		return 0;
	auto& p = c.pools[pool];
	kStagePool<<<blocksFor(occupancy, 256), 256, 0, c.stream>>>((const uint8_t*)dAos, stride, occupancy,
		p.aabbA, p.aabbB, p.entity, p.flags);
	return 1;
}

uint32_t launchLink(Context& c)
{
	uint32_t n = 0;
	for (uint32_t i = 0; i < c.poolCount; i++)
	{
		auto& p = c.pools[i];
		if (!p.set || p.occupancy == 0)
			continue;
		kLinkPool<<<blocksFor(p.occupancy, 256), 256, 0, c.stream>>>(p.occupancy, p.entity, c.tf.entityToSlot,
			c.tf.entityCap, p.tslot);
		n++;
	}
	return n;
}

} // namespace gsp
