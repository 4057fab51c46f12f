// Fused hot loop: world matrix (leaf-first parent-chain product) -> AABB-vs-frustum test for EVERY view of the frame ->
// distance key -> deterministic stream compaction (warp ballot + block scan + decoupled look-back across tiles).
//
// Replaces prepareUnsortedMeshes / prepareSortedMeshes (source/system/render/mesh.cpp:111-184,187-262) and what they call:
//   TransformComponent::calcModel        include/garden/system/transform.hpp:197-214
//   IMeshRenderSystem::getReadyMeshesAsync / isBehindFrustum   mesh.hpp:142-146, libraries/math/include/math/aabb.hpp:438-464
//   the key                              mesh.cpp:172,250-251
//   thread-local lists + fetch_add append mesh.cpp:177-183,257-261  (here: slot-ordered compaction, so ties sort by slot)
// The reference runs this loop once per view; here an entity's chain and corners are computed once and tested against
// all views (the camera position subtracted from the model is the same for every view, mesh.cpp:401,500).
#include "sceneprep_internal.h"
#include "sceneprep_math.cuh"
#include <string.h>

namespace gsp
{

struct CullArgs
{
	const float4* __restrict__ tRot;
	const float4* __restrict__ tPosSx;
	const float2* __restrict__ tSYZ;
	const uint32_t* __restrict__ tParent;
	const uint8_t* __restrict__ tFlags;
	const float4* __restrict__ aabbA;
	const float2* __restrict__ aabbB;
	const uint32_t* __restrict__ tslot;
	const uint8_t* __restrict__ mflags;
	const uint8_t* __restrict__ ready;
	float4* __restrict__ world;
	uint8_t* __restrict__ visible;
	uint32_t* __restrict__ status;   // [tiles][kMaxViews]
	uint32_t* __restrict__ counters;
	uint32_t* __restrict__ keys;
	uint32_t* __restrict__ payloads;
	uint32_t segOffset[kMaxViews];   // arena offset of this pool's list in view v
	uint32_t baseCounter[kMaxViews]; // counter index holding the list length before this pool (kNone = 0)
	uint32_t visibleView;            // view whose result is stored to isVisible (kNone = none)
	uint32_t tiles;
};

constexpr uint32_t kFlagAggregate = 1u << 30, kFlagInclusive = 2u << 30, kValueMask = (1u << 30) - 1;

__device__ __forceinline__ uint32_t ldVolatile(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void stRelease(uint32_t* p, uint32_t v)
{
	asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ Mat4 loadLocal(const CullArgs& a, uint32_t t)
{
	float4 q = a.tRot[t];
	float4 p = a.tPosSx[t];
	float2 s = a.tSYZ[t];
	return localModel(p.x, p.y, p.z, q.x, q.y, q.z, q.w, p.w, s.x, s.y);
}

__global__ void __launch_bounds__(kCullTile) kCull(const __grid_constant__ CullParams P, const __grid_constant__ CullArgs A)
{
	__shared__ uint32_t sTile;
	__shared__ uint32_t sWarp[kMaxViews][kCullTile / 32];
	__shared__ uint32_t sBase[kMaxViews];
	__shared__ uint32_t sInst[kMaxViews];

	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0)
		sTile = atomicAdd(&A.counters[kCtrCullTicket + P.poolIndex], 1u); // ticket => look-back cannot wait on an unscheduled tile
	if (threadIdx.x < kMaxViews)
		sInst[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t tile = sTile;
	const uint32_t slot = tile * kCullTile + threadIdx.x;

	// ---- filter (mesh.cpp:140-155) ----
	bool cand = slot < P.occupancy && (A.mflags[slot] & kMfCandidate);
	uint32_t ts = kNone;
	uint8_t tf = 0;
	if (cand) { ts = A.tslot[slot]; cand = ts != kNone; }
	if (cand) { tf = A.tFlags[ts]; cand = (tf & kTfLive) && (tf & kTfActive); }

	uint32_t mask = 0;
	uint32_t readyCount = 1;
	float c3x = 0.0f, c3y = 0.0f, c3z = 0.0f;
	if (cand)
	{
		// ---- world matrix: leaf-first chain product (transform.hpp:199-211) ----
		Mat4 M = loadLocal(A, ts);
		if (tf & kTfAncestors)
		{
			uint32_t p = A.tParent[ts];
			uint32_t depth = 0;
			while (p != kNone)
			{
				if (++depth > kMaxChainDepth)
				{
					atomicExch(&A.counters[kCtrError], (uint32_t)GSP_ERR_HIERARCHY);
					break;
				}
				Mat4 L = loadLocal(A, p);
				M = matMul(L, M);
				p = A.tParent[p];
			}
		}
		// translate(-cameraPosition, model): c3.xyz += -cam, w kept (matrix/transform.hpp:71-74)
		M.c[3][0] = __fadd_rn(M.c[3][0], -P.cam[0]);
		M.c[3][1] = __fadd_rn(M.c[3][1], -P.cam[1]);
		M.c[3][2] = __fadd_rn(M.c[3][2], -P.cam[2]);
		c3x = M.c[3][0]; c3y = M.c[3][1]; c3z = M.c[3][2];

		// ---- 8 corners (aabb.hpp:444-451) ----
		float4 ba = A.aabbA[slot];
		float2 bb = A.aabbB[slot];
		const float mn[3] = {ba.x, ba.y, ba.z}, mx[3] = {ba.w, bb.x, bb.y};
		float vx[8], vy[8], vz[8];
		#pragma unroll
		for (int k = 0; k < 8; k++)
			transformCorner(M, (k & 4) ? mx[0] : mn[0], (k & 2) ? mx[1] : mn[1], (k & 1) ? mx[2] : mn[2], vx[k], vy[k], vz[k]);

		// ---- plane tests per view (aabb.hpp:452-462): culled if some plane has all 8 corners at d < 0 ----
		for (uint32_t v = 0; v < P.viewCount; v++)
		{
			const ViewConst& V = P.views[v];
			if (!V.enabled)
				continue;
			bool culled = false;
			for (uint32_t i = 0; i < V.planeCount; i++)
			{
				const float nx = V.planes[i][0], ny = V.planes[i][1], nz = V.planes[i][2], nd = V.planes[i][3];
				bool allBehind = true;
				#pragma unroll
				for (int k = 0; k < 8; k++)
					allBehind = allBehind && (planeDistance(nx, ny, nz, nd, vx[k], vy[k], vz[k]) < 0.0f);
				if (allBehind) { culled = true; break; }
			}
			if (!culled)
				mask |= 1u << v;
		}
		if (P.hasReady) // a getReadyMeshesAsync override's extra predicate (e.g. sprite.cpp:90-97)
		{
			readyCount = A.ready[slot];
			if (readyCount == 0) mask = 0;
		}
		if (mask) // bakedModel = (float4x3)model (mesh.cpp:171,249)
		{
			float4* w = A.world + (size_t)slot * 3;
			w[0] = make_float4(M.c[0][0], M.c[0][1], M.c[0][2], M.c[1][0]);
			w[1] = make_float4(M.c[1][1], M.c[1][2], M.c[2][0], M.c[2][1]);
			w[2] = make_float4(M.c[2][2], M.c[3][0], M.c[3][1], M.c[3][2]);
		}
	}
	// isVisible of the (last) main view, written for every slot like mesh.cpp:144-146,152-153,161-167
	if (A.visibleView != kNone && slot < P.occupancy)
		A.visible[slot] = (uint8_t)((mask >> A.visibleView) & 1u);

	// ---- compaction: per view, warp ballot -> block scan -> decoupled look-back over tiles ----
	for (uint32_t v = 0; v < P.viewCount; v++)
	{
		uint32_t b = __ballot_sync(0xffffffffu, (mask >> v) & 1u);
		if (lane == 0)
			sWarp[v][warp] = __popc(b);
	}
	if (P.hasReady)
	{
		for (uint32_t v = 0; v < P.viewCount; v++)
			if ((mask >> v) & 1u)
				atomicAdd(&sInst[v], readyCount);
	}
	__syncthreads();
	if (warp == 0 && lane < P.viewCount)
	{
		const uint32_t v = lane;
		uint32_t total = 0;
		#pragma unroll
		for (uint32_t w = 0; w < kCullTile / 32; w++)
		{
			uint32_t cnt = sWarp[v][w];
			sWarp[v][w] = total;
			total += cnt;
		}
		uint32_t* st = A.status + (size_t)tile * kMaxViews + v;
		uint32_t exclusive = 0;
		if (tile == 0)
			stRelease(st, kFlagInclusive | total);
		else
		{
			stRelease(st, kFlagAggregate | total);
			int32_t t = (int32_t)tile - 1;
			while (true)
			{
				const uint32_t* ps = A.status + (size_t)t * kMaxViews + v;
				uint32_t s;
				do { s = ldVolatile(ps); } while ((s & ~kValueMask) == 0);
				exclusive += s & kValueMask;
				if (s & kFlagInclusive)
					break;
				t--;
			}
			__threadfence();
			stRelease(st, kFlagInclusive | (exclusive + total));
		}
		uint32_t listBase = A.baseCounter[v] != kNone ? A.counters[A.baseCounter[v]] : 0;
		sBase[v] = listBase + exclusive;
		if (tile == A.tiles - 1 && P.views[v].enabled)
		{
			A.counters[ctrPoolEnd(P.poolIndex, v)] = listBase + exclusive + total;
		}
		if (P.hasReady && sInst[v])
			atomicAdd(&A.counters[ctrPoolInst(P.poolIndex, v)], sInst[v]);
	}
	__syncthreads();
	// every lane takes part in the ballots (convergence), invisible lanes just skip the store
	{
		const uint32_t payload = (P.poolIndex << 28) | slot;
		for (uint32_t v = 0; v < P.viewCount; v++)
		{
			const bool vis = (mask >> v) & 1u;
			uint32_t b = __ballot_sync(0xffffffffu, vis);
			if (!vis)
				continue;
			uint32_t pos = sBase[v] + sWarp[v][warp] + __popc(b & ((1u << lane) - 1u));
			float key;
			if (P.key2D)
				key = __fadd_rn(c3z, 1.0f); // mesh.cpp:250
			else
			{
				const ViewConst& V = P.views[v];
				key = lengthSq3(__fadd_rn(c3x, V.cameraOffset[0]), __fadd_rn(c3y, V.cameraOffset[1]),
					__fadd_rn(c3z, V.cameraOffset[2])); // mesh.cpp:172,251
			}
			uint32_t k = floatToOrdered(key);
			if (P.descending)
				k = ~k;
			A.keys[A.segOffset[v] + pos] = k;
			A.payloads[A.segOffset[v] + pos] = payload;
		}
	}
}

uint32_t launchCull(Context& c, uint32_t pool)
{
	auto& p = c.pools[pool];
	if (!p.set || p.occupancy == 0)
		return 0;
	CullParams P = {};
	CullArgs A = {};
	const bool isUI = p.renderType == GSP_RT_UI;
	const bool sortedList = isUI || p.renderType == GSP_RT_TRANSLUCENT;
	bool any = false;
	A.visibleView = kNone;
	for (uint32_t v = 0; v < (uint32_t)c.views.size(); v++)
	{
		const gsp_view& gv = c.views[v];
		ViewConst& V = P.views[v];
		int seg = c.segOf[v][pool];
		V.enabled = seg >= 0 && c.participates[v][pool];
		A.segOffset[v] = 0; A.baseCounter[v] = kNone;
		if (!V.enabled)
			continue;
		any = true;
		memcpy(V.planes, isUI ? gv.uiPlanes : gv.planes, sizeof(V.planes));
		V.planeCount = isUI ? gv.uiPlaneCount : gv.planeCount;
		memcpy(V.cameraOffset, gv.cameraOffset, sizeof(V.cameraOffset));
		A.segOffset[v] = c.segments[seg].offset;
		int prev = c.prevPool[v][pool];
		A.baseCounter[v] = prev >= 0 ? ctrPoolEnd((uint32_t)prev, v) : kNone;
		if (gv.shadowPass < 0)
			A.visibleView = v; // the last main view wins, as repeated prepareMeshes calls would overwrite isVisible
	}
	if (!any)
		return 0;
	for (int i = 0; i < 3; i++)
		P.cam[i] = isUI ? 0.0f : c.cameraPos[i]; // mesh.cpp:435,441
	P.viewCount = (uint32_t)c.views.size();
	P.occupancy = p.occupancy;
	P.poolIndex = pool;
	P.key2D = isUI ? 1 : 0;
	P.descending = sortedList ? 1 : 0;
	P.hasReady = p.hasReady ? 1 : 0;

	A.tRot = c.tf.rot; A.tPosSx = c.tf.posSx; A.tSYZ = c.tf.sYZ; A.tParent = c.tf.parent; A.tFlags = c.tf.flags;
	A.aabbA = p.aabbA; A.aabbB = p.aabbB; A.tslot = p.tslot; A.mflags = p.flags; A.ready = p.ready;
	A.world = p.world; A.visible = p.visible;
	A.status = p.cullStatus; A.counters = c.dCounters;
	A.keys = c.keys[0]; A.payloads = c.payloads[0];
	A.tiles = (p.occupancy + kCullTile - 1) / kCullTile;
	p.visibleValid = A.visibleView != kNone;

	cudaMemsetAsync(p.cullStatus, 0, (size_t)A.tiles * kMaxViews * sizeof(uint32_t), c.stream);
	kCull<<<A.tiles, kCullTile, 0, c.stream>>>(P, A);
	return 1;
}

} // namespace gsp
