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
#include <math.h>
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
	uint32_t* __restrict__ visBits;  // [kMaxViews][tiles * 8] one ballot word per warp and view: visibility bit per slot
	uint32_t* __restrict__ tileCount; // [kMaxViews][tiles] visible slots per tile and view; scanned in place to list offsets
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

// ---- shared-memory cache of local matrices ----------------------------------------------------------------------------------
// Leaf-first association (transform.hpp:204-210) forces every entity to multiply its own chain, but the chain's factors —
// the ancestors' LOCAL matrices — are shared. Each tile computes the local matrix of every transform it touches once
// (bit-identical to recomputing it) and parks it in a direct-mapped cache keyed by transform slot;
// ancestors outside the tile (or evicted by a conflicting slot) are recomputed from the SoA streams.
constexpr uint32_t kCacheSize = 512, kCacheMask = kCacheSize - 1, kHalo = 16;

struct CullShared
{
	float L[12][kCacheSize];   // component-major: consecutive slots hit consecutive banks
	uint32_t tag[kCacheSize];  // transform slot held by the entry (kNone = empty)
	uint32_t par[kCacheSize];  // its parent slot
	uint32_t warpCount[kMaxViews][kCullTile / 32];
	uint32_t base[kMaxViews];
	uint32_t inst[kMaxViews];
	uint32_t total[kMaxViews];
	uint32_t tile;
	uint32_t minSlot;
};

__device__ __forceinline__ Mat43 loadLocal43(const CullArgs& a, uint32_t t)
{
	float4 q = a.tRot[t];
	float4 p = a.tPosSx[t];
	float2 s = a.tSYZ[t];
	return localModel43(p.x, p.y, p.z, q.x, q.y, q.z, q.w, p.w, s.x, s.y);
}

__device__ __forceinline__ void cacheInsert(CullShared& sh, uint32_t t, uint32_t parent, const Mat43& L)
{
	const uint32_t e = t & kCacheMask;
	// claim the entry first: two slots of one tile may collide, and only the winner may write the payload
	if (atomicCAS(&sh.tag[e], kNone, t) != kNone)
		return;
	#pragma unroll
	for (int i = 0; i < 4; i++)
		#pragma unroll
		for (int l = 0; l < 3; l++)
			sh.L[i * 3 + l][e] = L.c[i][l];
	sh.par[e] = parent;
}

// Conservative half-width of the band around a plane inside which the exact 8-corner test decides:
// computed plane distances differ from real arithmetic by a few ulp of the magnitudes involved (|n|_1 * A + |d|, A bounding
// every |corner lane| and every intermediate of the corner transform); 2^-16 of that leaves a factor > 30 of head room.
constexpr float kBandScale = 1.0f / 65536.0f;

__global__ void __launch_bounds__(kCullTile, 3) kCull(const __grid_constant__ CullParams P, const __grid_constant__ CullArgs A)
{
	__shared__ CullShared sh;

	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0)
		sh.minSlot = kNone;
	if (threadIdx.x < kMaxViews)
	{
		sh.inst[threadIdx.x] = 0;
		sh.total[threadIdx.x] = 0;
	}
	for (uint32_t i = threadIdx.x; i < kCacheSize; i += kCullTile)
		sh.tag[i] = kNone;
	__syncthreads();
	const uint32_t tile = blockIdx.x;
	const uint32_t slot = tile * kCullTile + threadIdx.x;

	// ---- filter (mesh.cpp:140-155) ----
	const bool inRange = slot < P.occupancy;
	bool cand = inRange && (A.mflags[slot] & kMfCandidate);
	uint32_t ts = inRange ? A.tslot[slot] : kNone;
	// flags, TRS and parent link depend only on `ts`: issue all the loads together (one latency, not four)
	uint8_t tf = 0;
	float4 tq = make_float4(0.f, 0.f, 0.f, 1.f), tp = make_float4(0.f, 0.f, 0.f, 1.f);
	float2 tsyz = make_float2(1.f, 1.f);
	uint32_t parentLink = kNone;
	if (ts != kNone)
	{
		tf = A.tFlags[ts]; tq = A.tRot[ts]; tp = A.tPosSx[ts]; tsyz = A.tSYZ[ts]; parentLink = A.tParent[ts];
	}
	const bool liveTransform = (tf & kTfLive) != 0;
	cand = cand && liveTransform && (tf & kTfActive);

	// ---- phase 1: local matrix of the own transform (and of a short halo before the tile) into the cache ----
	Mat43 M;
	uint32_t parent = kNone;
	if (liveTransform)
	{
		M = localModel43(tp.x, tp.y, tp.z, tq.x, tq.y, tq.z, tq.w, tp.w, tsyz.x, tsyz.y);
		parent = parentLink;
		cacheInsert(sh, ts, parent, M);
		atomicMin(&sh.minSlot, ts);
	}
	__syncthreads();
	if (threadIdx.x < kHalo)
	{
		// chains that start in the previous tile: their ancestors sit right before the tile's lowest transform slot
		const uint32_t first = sh.minSlot;
		if (first != kNone && first >= threadIdx.x + 1)
		{
			const uint32_t h = first - 1 - threadIdx.x;
			if (sh.tag[h & kCacheMask] == kNone && (A.tFlags[h] & kTfLive))  // (racing claims are settled by the CAS)
			{
				Mat43 H = loadLocal43(A, h);
				cacheInsert(sh, h, A.tParent[h], H);
			}
		}
	}
	__syncthreads();

	uint32_t mask = 0;
	uint32_t readyCount = 1;
	if (cand)
	{
		// ---- phase 2: world matrix, leaf-first chain product (transform.hpp:199-211) ----
		if (tf & kTfAncestors)
		{
			uint32_t p = parent;
			uint32_t depth = 0;
			while (p != kNone)
			{
				if (++depth > kMaxChainDepth)
				{
					atomicExch(&A.counters[kCtrError], (uint32_t)GSP_ERR_HIERARCHY);
					break;
				}
				const uint32_t e = p & kCacheMask;
				Mat43 L;
				uint32_t next;
				if (sh.tag[e] == p)
				{
					#pragma unroll
					for (int i = 0; i < 4; i++)
						#pragma unroll
						for (int l = 0; l < 3; l++)
							L.c[i][l] = sh.L[i * 3 + l][e];
					next = sh.par[e];
				}
				else
				{
					L = loadLocal43(A, p);
					next = A.tParent[p];
				}
				M = matMul43(L, M);
				p = next;
			}
		}
		// translate(-cameraPosition, model): c3.xyz += -cam, w kept (matrix/transform.hpp:71-74)
		M.c[3][0] = __fadd_rn(M.c[3][0], -P.cam[0]);
		M.c[3][1] = __fadd_rn(M.c[3][1], -P.cam[1]);
		M.c[3][2] = __fadd_rn(M.c[3][2], -P.cam[2]);

		const float4 ba = A.aabbA[slot];
		const float2 bb = A.aabbB[slot];
		const float mn[3] = {ba.x, ba.y, ba.z}, mx[3] = {ba.w, bb.x, bb.y};

		// ---- phase 3a: conservative bounds of the transformed box (any rounding is fine here, the band absorbs it) ----
		float ctr[3], ext[3], amax[3];
		#pragma unroll
		for (int k = 0; k < 3; k++)
		{
			ctr[k] = 0.5f * (mn[k] + mx[k]);
			ext[k] = 0.5f * fabsf(mx[k] - mn[k]);
			amax[k] = fmaxf(fabsf(mn[k]), fabsf(mx[k]));
		}
		float cw[3]; // world-space centre
		float magnitude = 0.0f; // A: bounds |lane| of every corner and of every partial sum of the corner transform
		#pragma unroll
		for (int l = 0; l < 3; l++)
		{
			cw[l] = fmaf(M.c[0][l], ctr[0], fmaf(M.c[1][l], ctr[1], fmaf(M.c[2][l], ctr[2], M.c[3][l])));
			float a = fmaf(fabsf(M.c[0][l]), amax[0], fmaf(fabsf(M.c[1][l]), amax[1], fmaf(fabsf(M.c[2][l]), amax[2], fabsf(M.c[3][l]))));
			magnitude = fmaxf(magnitude, a);
		}
		// radius of a sphere around the centre containing all corners: sum_i |c_i| * ext_i <= sqrt(3 * sum_i |c_i|^2 ext_i^2)
		float rr = 0.0f;
		#pragma unroll
		for (int i = 0; i < 3; i++)
		{
			float len2 = fmaf(M.c[i][0], M.c[i][0], fmaf(M.c[i][1], M.c[i][1], M.c[i][2] * M.c[i][2]));
			rr = fmaf(len2, ext[i] * ext[i], rr);
		}
		const float radius = sqrtf(3.0f * rr) * 1.0001f;

		// corners for the exact test are produced lazily, once
		float vx[8], vy[8], vz[8];
		bool haveCorners = false;

		// ---- phase 3b: plane tests per view (aabb.hpp:452-462): culled if some plane has all 8 corners at d < 0 ----
		for (uint32_t v = 0; v < P.viewCount; v++)
		{
			const ViewConst& V = P.views[v];
			if (!V.enabled)
				continue;
			bool culled = false;
			for (uint32_t i = 0; i < V.planeCount; i++)
			{
				const float nx = V.planes[i][0], ny = V.planes[i][1], nz = V.planes[i][2], nd = V.planes[i][3];
				const float dc = fmaf(nx, cw[0], fmaf(ny, cw[1], fmaf(nz, cw[2], nd)));
				const float band = fmaf(V.planeL1[i], magnitude, V.planeAbsD[i]); // already scaled by 2 * kBandScale
				const float reach = fmaf(radius, V.planeL2[i], band);
				if (dc > reach)
					continue;          // every corner is certainly in front: the plane cannot cull
				if (dc < -reach)
				{
					culled = true;     // every corner is certainly behind (all eight d < 0)
					break;
				}
				// straddling (or NaN): the reference's exact arithmetic decides
				if (!haveCorners)
				{
					#pragma unroll
					for (int k = 0; k < 8; k++)
						transformCorner43(M, (k & 4) ? mx[0] : mn[0], (k & 2) ? mx[1] : mn[1], (k & 1) ? mx[2] : mn[2],
							vx[k], vy[k], vz[k]);
					haveCorners = true;
				}
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
	if (A.visibleView != kNone && inRange)
		A.visible[slot] = (uint8_t)((mask >> A.visibleView) & 1u);

	// ---- visibility bits: one ballot word per warp and view, plus the tile's visible count per view ----
	// (no inter-tile dependency in this kernel: list positions are assigned by kScanTiles + kScatter below)
	for (uint32_t v = 0; v < P.viewCount; v++)
	{
		const uint32_t b = __ballot_sync(0xffffffffu, (mask >> v) & 1u);
		if (lane == 0)
		{
			A.visBits[((size_t)v * A.tiles + tile) * (kCullTile / 32) + warp] = b;
			if (b)
				atomicAdd(&sh.total[v], (uint32_t)__popc(b));
		}
	}
	if (P.hasReady)
	{
		for (uint32_t v = 0; v < P.viewCount; v++)
			if ((mask >> v) & 1u)
				atomicAdd(&sh.inst[v], readyCount);
	}
	__syncthreads();
	if (threadIdx.x < P.viewCount)
	{
		const uint32_t v = threadIdx.x;
		A.tileCount[(size_t)v * A.tiles + tile] = sh.total[v];
		if (P.hasReady && sh.inst[v])
			atomicAdd(&A.counters[ctrPoolInst(P.poolIndex, v)], sh.inst[v]);
	}
}

// Exclusive scan of the per-tile visible counts of one view (one block per view) -> list offset of every tile.
// Also publishes the list length after this pool (poolEnd), which is the draw count the host reads back.
constexpr uint32_t kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) kScanTiles(const __grid_constant__ CullParams P, const __grid_constant__ CullArgs A)
{
	const uint32_t v = blockIdx.x;
	if (!P.views[v].enabled)
		return;
	__shared__ uint32_t sWarp[kScanThreads / 32];
	uint32_t* counts = A.tileCount + (size_t)v * A.tiles;
	const uint32_t per = (A.tiles + kScanThreads - 1) / kScanThreads;
	const uint32_t begin = min(threadIdx.x * per, A.tiles), end = min(begin + per, A.tiles);
	uint32_t sum = 0;
	for (uint32_t i = begin; i < end; i++)
		sum += counts[i];
	// block exclusive scan of the per-thread sums
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = sum;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= (uint32_t)o) inc += t;
	}
	if (lane == 31)
		sWarp[warp] = inc;
	__syncthreads();
	if (warp == 0)
	{
		uint32_t w = sWarp[lane];
		uint32_t winc = w;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
			if (lane >= (uint32_t)o) winc += t;
		}
		sWarp[lane] = winc - w;
	}
	__syncthreads();
	const uint32_t listBase = A.baseCounter[v] != kNone ? A.counters[A.baseCounter[v]] : 0;
	uint32_t running = listBase + sWarp[warp] + inc - sum;
	for (uint32_t i = begin; i < end; i++)
	{
		const uint32_t c = counts[i];
		counts[i] = running;
		running += c;
	}
	if (threadIdx.x == kScanThreads - 1)
		A.counters[ctrPoolEnd(P.poolIndex, v)] = running;
}

// Compaction + key: every visible (slot, view) pair gets its position from the tile offset, the popcounts of the
// preceding warps' ballot words and its rank inside its own word -> lists come out in slot order (stable tie-break).
__global__ void __launch_bounds__(kCullTile) kScatter(const __grid_constant__ CullParams P, const __grid_constant__ CullArgs A)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t tile = blockIdx.x;
	const uint32_t slot = tile * kCullTile + threadIdx.x;
	const uint32_t payload = (P.poolIndex << 28) | slot;
	bool loaded = false;
	float c3x = 0.0f, c3y = 0.0f, c3z = 0.0f;
	for (uint32_t v = 0; v < P.viewCount; v++)
	{
		const ViewConst& V = P.views[v];
		if (!V.enabled)
			continue;
		uint32_t w = 0;
		if (lane < kCullTile / 32)
			w = A.visBits[((size_t)v * A.tiles + tile) * (kCullTile / 32) + lane];
		const uint32_t mine = __shfl_sync(0xffffffffu, w, warp);
		if (mine == 0)
			continue; // warp-uniform
		uint32_t before = lane < warp ? __popc(w) : 0;
		#pragma unroll
		for (int o = 4; o > 0; o >>= 1)
			before += __shfl_xor_sync(0xffffffffu, before, o);
		before = __shfl_sync(0xffffffffu, before, 0);
		if (!((mine >> lane) & 1u))
			continue;
		if (!loaded)
		{
			const float4 w2 = A.world[(size_t)slot * 3 + 2]; // (c2.z, c3.x, c3.y, c3.z) of the float4x3 world matrix
			c3x = w2.y; c3y = w2.z; c3z = w2.w;
			loaded = true;
		}
		const uint32_t pos = A.tileCount[(size_t)v * A.tiles + tile] + before + __popc(mine & ((1u << lane) - 1u));
		float key;
		if (P.key2D)
			key = __fadd_rn(c3z, 1.0f); // mesh.cpp:250
		else
			key = lengthSq3(__fadd_rn(c3x, V.cameraOffset[0]), __fadd_rn(c3y, V.cameraOffset[1]),
				__fadd_rn(c3z, V.cameraOffset[2])); // mesh.cpp:172,251
		uint32_t k = floatToOrdered(key);
		if (P.descending)
			k = ~k;
		A.keys[A.segOffset[v] + pos] = k;
		A.payloads[A.segOffset[v] + pos] = payload;
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
		for (uint32_t i = 0; i < 6; i++)
		{
			const float* pl = V.planes[i];
			// rounded up a little: these only widen the band in which the exact test is used
			V.planeL2[i] = sqrtf(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2]) * 1.0001f;
			V.planeL1[i] = (fabsf(pl[0]) + fabsf(pl[1]) + fabsf(pl[2])) * (2.0f * kBandScale);
			V.planeAbsD[i] = fabsf(pl[3]) * (2.0f * kBandScale);
		}
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
	A.visBits = p.visBits; A.tileCount = p.cullStatus; A.counters = c.dCounters;
	A.keys = c.keys[0]; A.payloads = c.payloads[0];
	A.tiles = (p.occupancy + kCullTile - 1) / kCullTile;
	p.visibleValid = A.visibleView != kNone;

	kCull<<<A.tiles, kCullTile, 0, c.stream>>>(P, A);
	kScanTiles<<<P.viewCount, kScanThreads, 0, c.stream>>>(P, A);
	kScatter<<<A.tiles, kCullTile, 0, c.stream>>>(P, A);
	return 3;
}

} // namespace gsp
