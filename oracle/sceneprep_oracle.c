/* TEST INFRASTRUCTURE ONLY — see sceneprep_oracle.h.
 *
 * Plain scalar C restatement of the reference's per-frame scene preparation, in the reference's exact
 * floating-point operation order for its x86 AVX2 build without compiler contraction ("dialect B", SURVEY.md
 * finding 3): fmaf() exactly where the reference uses MATH_SIMD_FMA, every other operation separately rounded.
 * Must be compiled with -ffp-contract=off (oracle/Makefile does).
 *
 * Citations are to files under /root/reference.
 */
#include "sceneprep_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* TransformComponent field offsets, include/garden/system/transform.hpp:31-60 (80 bytes, Release non-editor). */
enum { T_ENTITY = 0, T_PARENT = 4, T_POS = 16, T_SCALE = 32, T_ROT = 48, T_SELF_ACTIVE = 72, T_ANC_ACTIVE = 73,
	T_WITH_ANCESTORS = 74 };
/* MeshRenderComponent field offsets, include/garden/system/render/mesh.hpp:45-55 (48 bytes). */
enum { M_ENTITY = 0, M_ENABLED = 14, M_VISIBLE = 15, M_AABB_MIN = 16, M_AABB_MAX = 32 };

typedef struct
{
	uint8_t* data;
	uint32_t stride, occupancy, count, renderType, drawReady, drawReadyShadow;
	const uint8_t* readyCounts;
	uint32_t readyCountsSize;
	uint8_t* visible;
	uint32_t visibleCap;
} Pool;

typedef struct
{
	OracleRecord* records;
	uint32_t cap, drawCount, instanceCount, pool;
} Buffer;

struct OracleScene
{
	const uint8_t* transforms;
	uint32_t tStride, tOccupancy;
	uint32_t* entityToSlot; /* entity id -> transform slot + 1 (0 = none) */
	uint32_t entityCap;
	Pool pools[ORACLE_MAX_POOLS];
	uint32_t poolCount;
	float cameraPos[3];
	Buffer unsorted[ORACLE_MAX_POOLS];
	uint32_t unsortedBufferCount, sortedBufferCount;
	uint32_t sortedDraw[ORACLE_MAX_POOLS], sortedInst[ORACLE_MAX_POOLS];
	Buffer trans, ui;
};

static uint32_t ld_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static float ld_f32(const uint8_t* p) { float v; memcpy(&v, p, 4); return v; }

/* f32x4x4::operator*(f32x4x4), libraries/math/include/math/simd/matrix/float.hpp:193-204:
 * per column of b: r = a.c0 * b.x; r = FMA(a.c1, b.y, r); r = FMA(a.c2, b.z, r); r = FMA(a.c3, b.w, r); all 4 lanes. */
void oracle_mat_mul(const float a[16], const float b[16], float out[16])
{
	float r[16];
	for (int i = 0; i < 4; i++)
	{
		for (int l = 0; l < 4; l++)
		{
			float v = a[l] * b[i * 4];
			v = fmaf(a[4 + l], b[i * 4 + 1], v);
			v = fmaf(a[8 + l], b[i * 4 + 2], v);
			v = fmaf(a[12 + l], b[i * 4 + 3], v);
			r[i * 4 + l] = v;
		}
	}
	memcpy(out, r, sizeof(r));
}

/* f32x4x4::operator*(f32x4), simd/matrix/float.hpp:225-231. */
static void mat_vec(const float m[16], float x, float y, float z, float w, float out[4])
{
	for (int l = 0; l < 4; l++)
	{
		float v = m[l] * x;
		v = fmaf(m[4 + l], y, v);
		v = fmaf(m[8 + l], z, v);
		v = fmaf(m[12 + l], w, v);
		out[l] = v;
	}
}

/* math::calcModel general branch, libraries/math/include/math/matrix/transform.hpp:255:
 * translate(position) * rotate(normalize(rotation)) * scale(scale). */
void oracle_local_model(const float pos[3], const float rot[4], const float scale[3], float out[16])
{
	/* normalize4, simd/vector/float.hpp:1198-1201: dpps 0xff = (x*x + y*y) + (z*z + w*w), sqrt, per-lane divide. */
	float d = (rot[0] * rot[0] + rot[1] * rot[1]) + (rot[2] * rot[2] + rot[3] * rot[3]);
	float n = sqrtf(d);
	float x = rot[0] / n, y = rot[1] / n, z = rot[2] / n, w = rot[3] / n;

	/* rotate(quat), matrix/transform.hpp:128-140 */
	float xx = x * x, yy = y * y, zz = z * z;
	float xz = x * z, xy = x * y, yz = y * z;
	float wx = w * x, wy = w * y, wz = w * z;
	float R[16] = {
		1.0f - 2.0f * (yy + zz), 2.0f * (xy + wz), 2.0f * (xz - wy), 0.0f,
		2.0f * (xy - wz), 1.0f - 2.0f * (xx + zz), 2.0f * (yz + wx), 0.0f,
		2.0f * (xz + wy), 2.0f * (yz - wx), 1.0f - 2.0f * (xx + yy), 0.0f,
		0.0f, 0.0f, 0.0f, 1.0f };
	/* translate(t), matrix/transform.hpp:50-54; scale(s), :80-84 */
	float T[16] = { 1, 0, 0, 0,  0, 1, 0, 0,  0, 0, 1, 0,  pos[0], pos[1], pos[2], 1.0f };
	float S[16] = { scale[0], 0, 0, 0,  0, scale[1], 0, 0,  0, 0, scale[2], 0,  0, 0, 0, 1.0f };
	float TR[16];
	oracle_mat_mul(T, R, TR);
	oracle_mat_mul(TR, S, out);
}

/* math::calcModel including its `scale == f32x4::one` branch (matrix/transform.hpp:251-256). The comparison is over
 * all four lanes and lane W of scaleChildCap holds childCapacity bits (transform.hpp:40,52), so for components the
 * branch is only taken if those bits equal 1.0f; kept for fidelity. */
static void local_model_of(const uint8_t* t, float out[16])
{
	float pos[3], rot[4], scale[3];
	for (int i = 0; i < 3; i++) { pos[i] = ld_f32(t + T_POS + 4 * i); scale[i] = ld_f32(t + T_SCALE + 4 * i); }
	for (int i = 0; i < 4; i++) rot[i] = ld_f32(t + T_ROT + 4 * i);
	float scaleW = ld_f32(t + T_SCALE + 12);
	if (scale[0] == 1.0f && scale[1] == 1.0f && scale[2] == 1.0f && scaleW == 1.0f)
	{
		/* translate(position, rotate(normalize(q))): c3 = (R.c3 + position).xyz, w kept (:71-74) */
		float one[3] = { 1.0f, 1.0f, 1.0f }, zero[3] = { 0.0f, 0.0f, 0.0f };
		float d = (rot[0] * rot[0] + rot[1] * rot[1]) + (rot[2] * rot[2] + rot[3] * rot[3]);
		float n = sqrtf(d);
		float x = rot[0] / n, y = rot[1] / n, z = rot[2] / n, w = rot[3] / n;
		float xx = x * x, yy = y * y, zz = z * z, xz = x * z, xy = x * y, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
		float R[16] = {
			1.0f - 2.0f * (yy + zz), 2.0f * (xy + wz), 2.0f * (xz - wy), 0.0f,
			2.0f * (xy - wz), 1.0f - 2.0f * (xx + zz), 2.0f * (yz + wx), 0.0f,
			2.0f * (xz + wy), 2.0f * (yz - wx), 1.0f - 2.0f * (xx + yy), 0.0f,
			0.0f + pos[0], 0.0f + pos[1], 0.0f + pos[2], 1.0f };
		(void)one; (void)zero;
		memcpy(out, R, sizeof(R));
		return;
	}
	oracle_local_model(pos, rot, scale, out);
}

static const uint8_t* transform_of_entity(const OracleScene* s, uint32_t entity)
{
	if (entity == 0 || entity >= s->entityCap)
		return NULL;
	uint32_t slot1 = s->entityToSlot[entity];
	return slot1 ? s->transforms + (size_t)(slot1 - 1) * s->tStride : NULL;
}

/* TransformComponent::calcModel, include/garden/system/transform.hpp:197-214: leaf-first chain product,
 * then translate(-cameraPosition, model) (matrix/transform.hpp:71-74: one add on c3.xyz, w kept). */
static int calc_model(const OracleScene* s, const uint8_t* t, const float cam[3], float model[16])
{
	local_model_of(t, model);
	if (t[T_WITH_ANCESTORS])
	{
		uint32_t nextParent = ld_u32(t + T_PARENT);
		uint32_t guard = 0;
		while (nextParent)
		{
			const uint8_t* p = transform_of_entity(s, nextParent);
			if (!p || ++guard > s->tOccupancy)
				return -1; /* reference: manager->get<> throws */
			float parentModel[16];
			local_model_of(p, parentModel);
			oracle_mat_mul(parentModel, model, model);
			nextParent = ld_u32(p + T_PARENT);
		}
	}
	for (int i = 0; i < 3; i++)
		model[12 + i] = model[12 + i] + (-cam[i]);
	return 0;
}

/* dot3 = dpps 0x7f, simd/vector/float.hpp:1090-1093: (a.x*b.x + a.y*b.y) + (a.z*b.z + 0) */
static float dot3(const float a[4], const float b[4])
{
	return (a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + 0.0f);
}

/* isBehindFrustum, libraries/math/include/math/aabb.hpp:438-464; distance3, plane.hpp:115-118. */
static int is_behind_frustum(const float (*planes)[4], uint32_t planeCount, const float mn[3], const float mx[3],
	const float model[16])
{
	float v[8][4];
	mat_vec(model, mn[0], mn[1], mn[2], 1.0f, v[0]);
	mat_vec(model, mn[0], mn[1], mx[2], 1.0f, v[1]);
	mat_vec(model, mn[0], mx[1], mn[2], 1.0f, v[2]);
	mat_vec(model, mn[0], mx[1], mx[2], 1.0f, v[3]);
	mat_vec(model, mx[0], mn[1], mn[2], 1.0f, v[4]);
	mat_vec(model, mx[0], mn[1], mx[2], 1.0f, v[5]);
	mat_vec(model, mx[0], mx[1], mn[2], 1.0f, v[6]);
	mat_vec(model, mx[0], mx[1], mx[2], 1.0f, v[7]);
	for (uint32_t i = 0; i < planeCount; i++)
	{
		int all = 1;
		for (int k = 0; k < 8; k++)
		{
			float d = dot3(planes[i], v[k]) + planes[i][3];
			if (!(d < 0.0f))
				all = 0;
		}
		if (all)
			return 1;
	}
	return 0;
}

static void buffer_reserve(Buffer* b, uint32_t cap)
{
	if (b->cap < cap)
	{
		b->records = (OracleRecord*)realloc(b->records, (size_t)cap * sizeof(OracleRecord));
		b->cap = cap;
	}
}

/* prepareUnsortedMeshes / prepareSortedMeshes, source/system/render/mesh.cpp:111-184,187-262 (single-threaded order). */
static int prepare_pool(OracleScene* s, Pool* pool, const float (*planes)[4], uint32_t planeCount,
	const float cam[3], const float cameraOffset[4], int isNotShadowPass, int distance2D, uint32_t bufferIndex,
	Buffer* out, uint32_t* drawCountOut, uint32_t* instanceCountOut, int writeVisible)
{
	uint32_t drawCount = 0, instanceCount = 0;
	for (uint32_t i = 0; i < pool->occupancy; i++)
	{
		uint8_t* m = pool->data + (size_t)i * pool->stride;
		float mn[3], mx[3];
		for (int k = 0; k < 3; k++) { mn[k] = ld_f32(m + M_AABB_MIN + 4 * k); mx[k] = ld_f32(m + M_AABB_MAX + 4 * k); }
		int visible = 0;
		uint32_t entity = ld_u32(m + M_ENTITY);
		/* mesh.cpp:140-147: getSize() = max - min, fixW() (w = z), areAllTrue(size <= 0) */
		int degenerate = ((mx[0] - mn[0]) <= 0.0f) && ((mx[1] - mn[1]) <= 0.0f) && ((mx[2] - mn[2]) <= 0.0f);
		if (entity && m[M_ENABLED] && !degenerate)
		{
			const uint8_t* t = transform_of_entity(s, entity); /* mesh.cpp:149-155 */
			if (t && t[T_SELF_ACTIVE] && t[T_ANC_ACTIVE])
			{
				float model[16];
				if (calc_model(s, t, cam, model) != 0)
					return -1;
				uint32_t readyCount = is_behind_frustum(planes, planeCount, mn, mx, model) ? 0 : 1; /* mesh.hpp:142-146 */
				if (readyCount && pool->readyCounts)
					readyCount = i < pool->readyCountsSize ? pool->readyCounts[i] : 1;
				if (readyCount)
				{
					visible = 1;
					OracleRecord* r = &out->records[out->drawCount + drawCount]; /* mesh.cpp:169-173 / 247-253 */
					memset(r, 0, sizeof(*r));
					r->componentOffset = (uint64_t)i * pool->stride;
					for (int c = 0; c < 4; c++)
						for (int l = 0; l < 3; l++)
							r->bakedModel[c * 3 + l] = model[c * 4 + l];
					if (distance2D)
						r->distanceSq = model[14] + 1.0f; /* mesh.cpp:250 */
					else
					{
						float u[4];
						for (int l = 0; l < 4; l++) u[l] = model[12 + l] + cameraOffset[l];
						r->distanceSq = dot3(u, u); /* lengthSq3, simd/vector/float.hpp:1148 */
					}
					r->bufferIndex = bufferIndex;
					drawCount++;
					instanceCount += readyCount;
				}
			}
		}
		if (isNotShadowPass)
		{
			pool->visible[i] = (uint8_t)visible;
			if (writeVisible)
				m[M_VISIBLE] = (uint8_t)visible;
		}
	}
	out->drawCount += drawCount;
	*drawCountOut += drawCount;
	*instanceCountOut += instanceCount;
	return 0;
}

/* Canonical order: the reference's std::sort on distanceSq only (mesh.hpp:196,204; mesh.cpp:285,303,319) leaves tie order
 * unspecified; the canonical tie-break is (bufferIndex, componentOffset) ascending. */
static int cmp_ascending(const void* pa, const void* pb)
{
	const OracleRecord* a = (const OracleRecord*)pa; const OracleRecord* b = (const OracleRecord*)pb;
	if (a->distanceSq < b->distanceSq) return -1;
	if (b->distanceSq < a->distanceSq) return 1;
	if (a->bufferIndex != b->bufferIndex) return a->bufferIndex < b->bufferIndex ? -1 : 1;
	if (a->componentOffset != b->componentOffset) return a->componentOffset < b->componentOffset ? -1 : 1;
	return 0;
}
static int cmp_descending(const void* pa, const void* pb)
{
	const OracleRecord* a = (const OracleRecord*)pa; const OracleRecord* b = (const OracleRecord*)pb;
	if (a->distanceSq > b->distanceSq) return -1;
	if (b->distanceSq > a->distanceSq) return 1;
	if (a->bufferIndex != b->bufferIndex) return a->bufferIndex < b->bufferIndex ? -1 : 1;
	if (a->componentOffset != b->componentOffset) return a->componentOffset < b->componentOffset ? -1 : 1;
	return 0;
}

OracleScene* oracle_create(void) { return (OracleScene*)calloc(1, sizeof(OracleScene)); }

void oracle_destroy(OracleScene* s)
{
	if (!s) return;
	free(s->entityToSlot);
	for (int i = 0; i < ORACLE_MAX_POOLS; i++) { free(s->pools[i].visible); free(s->unsorted[i].records); }
	free(s->trans.records); free(s->ui.records);
	free(s);
}

/* Entity -> TransformComponent lookup. The reference resolves it through the entity's component list
 * (Manager::tryGet, libraries/ecsm/include/ecsm.hpp:898-905); every live TransformComponent stores its owner
 * entity at offset 0, so the inverse map over the pool is equivalent. */
int oracle_set_transforms(OracleScene* s, const void* data, uint32_t stride, uint32_t occupancy)
{
	s->transforms = (const uint8_t*)data; s->tStride = stride; s->tOccupancy = occupancy;
	uint32_t maxEntity = 0;
	for (uint32_t i = 0; i < occupancy; i++)
	{
		uint32_t e = ld_u32(s->transforms + (size_t)i * stride + T_ENTITY);
		if (e > maxEntity) maxEntity = e;
	}
	free(s->entityToSlot);
	s->entityCap = maxEntity + 1;
	s->entityToSlot = (uint32_t*)calloc(s->entityCap, sizeof(uint32_t));
	for (uint32_t i = 0; i < occupancy; i++)
	{
		uint32_t e = ld_u32(s->transforms + (size_t)i * stride + T_ENTITY);
		if (e) s->entityToSlot[e] = i + 1;
	}
	return 0;
}

int oracle_set_pool(OracleScene* s, uint32_t index, uint32_t renderType, uint32_t drawReady, void* data,
	uint32_t stride, uint32_t occupancy, uint32_t count, const uint8_t* readyCounts, uint32_t readyCountsSize)
{
	if (index >= ORACLE_MAX_POOLS) return -1;
	Pool* p = &s->pools[index];
	p->data = (uint8_t*)data; p->stride = stride; p->occupancy = occupancy; p->count = count;
	p->renderType = renderType; p->drawReady = p->drawReadyShadow = drawReady;
	p->readyCounts = readyCountsSize ? readyCounts : NULL; p->readyCountsSize = readyCountsSize;
	if (p->visibleCap < occupancy)
	{
		p->visible = (uint8_t*)realloc(p->visible, occupancy ? occupancy : 1);
		p->visibleCap = occupancy;
	}
	if (index >= s->poolCount) s->poolCount = index + 1;
	return 0;
}
void oracle_set_pool_count(OracleScene* s, uint32_t poolCount) { s->poolCount = poolCount; }
void oracle_set_pool_draw_ready(OracleScene* s, uint32_t pool, uint32_t readyMain, uint32_t readyShadow)
{
	if (pool >= ORACLE_MAX_POOLS) return;
	s->pools[pool].drawReady = readyMain; s->pools[pool].drawReadyShadow = readyShadow;
}
void oracle_set_camera(OracleScene* s, const float cameraPos[3]) { memcpy(s->cameraPos, cameraPos, 12); }

/* MeshRenderSystem::prepareMeshes, source/system/render/mesh.cpp:331-553. */
int oracle_prepare(OracleScene* s, const OracleView* view, int writeVisible)
{
	int isNotShadowPass = view->shadowPass < 0;
	uint32_t transMax = 0, uiMax = 0;
	s->unsortedBufferCount = s->sortedBufferCount = 0;
	s->trans.drawCount = s->ui.drawCount = 0;
	for (uint32_t i = 0; i < s->poolCount; i++) /* mesh.cpp:341-375 */
	{
		Pool* p = &s->pools[i];
		if (p->renderType == ORACLE_RT_TRANSLUCENT) { transMax += p->count; s->sortedBufferCount++; }
		else if (p->renderType == ORACLE_RT_UI) { if (isNotShadowPass) { uiMax += p->count; s->sortedBufferCount++; } }
		else s->unsortedBufferCount++;
	}
	buffer_reserve(&s->trans, transMax); buffer_reserve(&s->ui, uiMax);

	uint32_t unsortedIndex = 0, sortedIndex = 0;
	const float zero[3] = { 0.0f, 0.0f, 0.0f };
	for (uint32_t i = 0; i < s->poolCount; i++) /* mesh.cpp:408-523 */
	{
		Pool* p = &s->pools[i];
		if (isNotShadowPass) memset(p->visible, 0xFF, p->occupancy);
		if (p->renderType == ORACLE_RT_TRANSLUCENT || p->renderType == ORACLE_RT_UI)
		{
			if (p->renderType == ORACLE_RT_UI && !isNotShadowPass) continue;
			uint32_t bufferIndex = sortedIndex++;
			s->sortedDraw[bufferIndex] = s->sortedInst[bufferIndex] = 0;
			if (p->count == 0 || !(isNotShadowPass ? p->drawReady : p->drawReadyShadow)) continue; /* isDrawReady(shadowPass), mesh.cpp:426,482 */
			int rc;
			if (p->renderType == ORACLE_RT_TRANSLUCENT)
				rc = prepare_pool(s, p, view->planes, view->planeCount, s->cameraPos, view->cameraOffset, isNotShadowPass,
					0, bufferIndex, &s->trans, &s->sortedDraw[bufferIndex], &s->sortedInst[bufferIndex], writeVisible);
			else
			{
				if (view->uiPlaneCount == 0) return -2; /* reference would dereference a null uiFrustum */
				rc = prepare_pool(s, p, view->uiPlanes, view->uiPlaneCount, zero, view->cameraOffset, isNotShadowPass,
					1, bufferIndex, &s->ui, &s->sortedDraw[bufferIndex], &s->sortedInst[bufferIndex], writeVisible);
			}
			if (rc) return rc;
		}
		else
		{
			Buffer* b = &s->unsorted[unsortedIndex++];
			b->drawCount = b->instanceCount = 0; b->pool = i;
			if (p->count == 0 || !(isNotShadowPass ? p->drawReady : p->drawReadyShadow)) continue; /* isDrawReady(shadowPass), mesh.cpp:426,482 */
			buffer_reserve(b, p->occupancy);
			uint32_t draw = 0;
			int rc = prepare_pool(s, p, view->planes, view->planeCount, s->cameraPos, view->cameraOffset, isNotShadowPass,
				0, 0, b, &draw, &b->instanceCount, writeVisible);
			if (rc) return rc;
		}
	}

	/* sortMeshes, mesh.cpp:265-328 */
	for (uint32_t i = 0; i < s->unsortedBufferCount; i++)
	{
		Buffer* b = &s->unsorted[i];
		if (s->pools[b->pool].renderType == ORACLE_RT_OIT || b->drawCount == 0) continue;
		qsort(b->records, b->drawCount, sizeof(OracleRecord), cmp_ascending);
	}
	if (s->trans.drawCount) qsort(s->trans.records, s->trans.drawCount, sizeof(OracleRecord), cmp_descending);
	if (s->ui.drawCount) qsort(s->ui.records, s->ui.drawCount, sizeof(OracleRecord), cmp_descending);
	return 0;
}

uint32_t oracle_unsorted_buffer_count(const OracleScene* s) { return s->unsortedBufferCount; }
uint32_t oracle_sorted_buffer_count(const OracleScene* s) { return s->sortedBufferCount; }
void oracle_get_unsorted(const OracleScene* s, uint32_t buffer, const OracleRecord** records,
	uint32_t* drawCount, uint32_t* instanceCount)
{
	*records = s->unsorted[buffer].records; *drawCount = s->unsorted[buffer].drawCount;
	*instanceCount = s->unsorted[buffer].instanceCount;
}
void oracle_get_sorted_counts(const OracleScene* s, uint32_t buffer, uint32_t* drawCount, uint32_t* instanceCount)
{
	*drawCount = s->sortedDraw[buffer]; *instanceCount = s->sortedInst[buffer];
}
void oracle_get_sorted(const OracleScene* s, int which, const OracleRecord** records, uint32_t* drawCount)
{
	const Buffer* b = which == 0 ? &s->trans : &s->ui;
	*records = b->records; *drawCount = b->drawCount;
}
const uint8_t* oracle_get_visible(const OracleScene* s, uint32_t pool) { return s->pools[pool].visible; }

int oracle_calc_model(const OracleScene* s, uint32_t transformSlot, const float cameraPos[3], float out[16])
{
	if (transformSlot >= s->tOccupancy) return -1;
	return calc_model(s, s->transforms + (size_t)transformSlot * s->tStride, cameraPos, out);
}

/* Frustum(viewProj), libraries/math/include/math/frustum.hpp:51-61 (Gribb-Hartmann, Vulkan Y flip). */
void oracle_frustum_planes(const float m[16], float planes[6][4])
{
	/* t = transpose4x4(viewProj): t.c_i lane j = m.c_j lane i */
	float t[4][4];
	for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t[i][j] = m[j * 4 + i];
	for (int l = 0; l < 4; l++)
	{
		planes[0][l] = t[3][l] + t[0][l];
		planes[1][l] = t[3][l] - t[0][l];
		planes[2][l] = t[3][l] - t[1][l];
		planes[3][l] = t[3][l] + t[1][l];
		planes[4][l] = t[2][l];
		planes[5][l] = t[3][l] - t[2][l];
	}
}

/* ---- SURVEY.md §8f rows ----------------------------------------------------------------------------------------------- */

/* f1. renderUnsorted / renderSorted rebuild the model as f32x4x4(mesh.bakedModel, f32x4(0, 0, 0, 1)) (mesh.cpp:600,632,
 * simd/matrix/float.hpp:87-89: lanes xyz from the float4x3 columns, lane W of column i = r3[i]) and every drawAsync
 * stores mvp = (float4x4)(viewProj * model) into the instance buffer (sprite.cpp:122-130). */
void oracle_instance_mvp(const float viewProj[16], const OracleRecord* records, uint32_t count, float* out)
{
	for (uint32_t i = 0; i < count; i++)
	{
		float model[16];
		for (int c = 0; c < 4; c++)
		{
			model[c * 4 + 0] = records[i].bakedModel[c * 3 + 0];
			model[c * 4 + 1] = records[i].bakedModel[c * 3 + 1];
			model[c * 4 + 2] = records[i].bakedModel[c * 3 + 2];
			model[c * 4 + 3] = c == 3 ? 1.0f : 0.0f;
		}
		oracle_mat_mul(viewProj, model, out + (size_t)i * 16);
	}
}

/* f3. TransformComponent::setActive, source/system/transform.cpp:75-127, on the AoS pool. The reference walks the
 * `childs` arrays with an explicit stack; the arrays are not part of the pool bytes, so child lists are rebuilt here from
 * the parent links (counting sort by parent), which visits the same set of descendants. */
int oracle_set_active(void* transforms, uint32_t stride, uint32_t occupancy, const uint32_t* entityIds, uint32_t count,
	int active)
{
	uint8_t* base = (uint8_t*)transforms;
	if (!base || stride < 80 || (!entityIds && count))
		return -1;
	uint32_t maxEntity = 0;
	for (uint32_t i = 0; i < occupancy; i++)
	{
		uint32_t e = ld_u32(base + (size_t)i * stride + T_ENTITY);
		if (e > maxEntity) maxEntity = e;
	}
	/* entity id -> slot + 1 (Manager::get<TransformComponent>(entity)) */
	uint32_t* slotOf = (uint32_t*)calloc((size_t)maxEntity + 1, sizeof(uint32_t));
	uint32_t* childStart = (uint32_t*)calloc((size_t)occupancy + 2, sizeof(uint32_t));
	uint32_t* childList = (uint32_t*)malloc(((size_t)occupancy + 1) * sizeof(uint32_t));
	uint32_t* stack = (uint32_t*)malloc(((size_t)occupancy + 1) * sizeof(uint32_t));
	if (!slotOf || !childStart || !childList || !stack)
	{
		free(slotOf); free(childStart); free(childList); free(stack);
		return -1;
	}
	for (uint32_t i = 0; i < occupancy; i++)
	{
		uint32_t e = ld_u32(base + (size_t)i * stride + T_ENTITY);
		if (e) slotOf[e] = i + 1;
	}
	for (uint32_t i = 0; i < occupancy; i++)
	{
		const uint8_t* t = base + (size_t)i * stride;
		uint32_t e = ld_u32(t + T_ENTITY), p = ld_u32(t + T_PARENT);
		if (e && p && p <= maxEntity && slotOf[p])
			childStart[slotOf[p] - 1 + 1]++;
	}
	for (uint32_t i = 0; i < occupancy; i++)
		childStart[i + 1] += childStart[i];
	{
		uint32_t* fill = (uint32_t*)calloc((size_t)occupancy + 1, sizeof(uint32_t));
		if (!fill) { free(slotOf); free(childStart); free(childList); free(stack); return -1; }
		for (uint32_t i = 0; i < occupancy; i++)
		{
			const uint8_t* t = base + (size_t)i * stride;
			uint32_t e = ld_u32(t + T_ENTITY), p = ld_u32(t + T_PARENT);
			if (e && p && p <= maxEntity && slotOf[p])
			{
				uint32_t ps = slotOf[p] - 1;
				childList[childStart[ps] + fill[ps]++] = i;
			}
		}
		free(fill);
	}
	int rc = 0;
	for (uint32_t k = 0; k < count; k++)
	{
		uint32_t e = entityIds[k];
		if (!e || e > maxEntity || !slotOf[e]) { rc = -1; continue; }
		uint8_t* self = base + (size_t)(slotOf[e] - 1) * stride;
		uint8_t isActive = active ? 1 : 0;
		if ((self[T_SELF_ACTIVE] != 0) == (isActive != 0))
			continue;                               /* :77-78 */
		self[T_SELF_ACTIVE] = isActive;             /* :80 */
		uint32_t top = 0;
		stack[top++] = slotOf[e] - 1;               /* :83 */
		if (isActive)
		{
			if (!self[T_ANC_ACTIVE])
				continue;                           /* :87-88 */
			while (top)
			{
				uint32_t s = stack[--top];
				uint8_t* t = base + (size_t)s * stride;
				if (!t[T_SELF_ACTIVE])
					continue;                       /* :94-95 */
				for (uint32_t c = childStart[s]; c < childStart[s + 1]; c++)
				{
					base[(size_t)childList[c] * stride + T_ANC_ACTIVE] = 1; /* :104 */
					stack[top++] = childList[c];
				}
			}
		}
		else
		{
			while (top)
			{
				uint32_t s = stack[--top];
				for (uint32_t c = childStart[s]; c < childStart[s + 1]; c++)
				{
					base[(size_t)childList[c] * stride + T_ANC_ACTIVE] = 0; /* :122 */
					stack[top++] = childList[c];
				}
			}
		}
	}
	free(slotOf); free(childStart); free(childList); free(stack);
	return rc;
}

/* f2. TransformSystem::animateAsync, source/system/transform.cpp:609-623.
 * lerp(f32x4 a, f32x4 b, float t) = a * (1.0f - t) + b * t (simd/vector/float.hpp:1469): two lane-wise muls and an add, no FMA
 * (intrinsics, dialect B). slerp (quaternion.hpp:175-193): dot4 (dpps 0xff: (x*x + y*y) + (z*z + w*w)), shortest path by
 * negating b, lerp when cosTheta > 1 - FLT_EPSILON, else (a * sin((1 - t) * angle) + c * sin(t * angle)) / sin(angle) with
 * std::acos / std::sin on floats = acosf / sinf. setPosition / setScale keep lane W (childCount / childCapacity bits). */
static void lerp3(const float* a, const float* b, float t, float* out)
{
	const float s = 1.0f - t;
	for (int l = 0; l < 3; l++)
		out[l] = a[l] * s + b[l] * t;
}
int oracle_animate(void* transforms, uint32_t stride, uint32_t occupancy, uint32_t count, const uint32_t* entityIds,
	const uint8_t* flags, const float* frameA, const float* frameB, const float* t)
{
	uint8_t* base = (uint8_t*)transforms;
	if (!base || stride < 80)
		return -1;
	int rc = 0;
	for (uint32_t i = 0; i < count; i++)
	{
		uint8_t* self = NULL;
		for (uint32_t sl = 0; sl < occupancy; sl++) /* Manager::get<TransformComponent>(entity) */
			if (ld_u32(base + (size_t)sl * stride + T_ENTITY) == entityIds[i] && entityIds[i])
			{
				self = base + (size_t)sl * stride;
				break;
			}
		if (!self) { rc = -1; continue; }
		const float* a = frameA + (size_t)i * 10; const float* b = frameB + (size_t)i * 10;
		const float ti = t[i];
		float v[4];
		if (flags[i] & 1)
		{
			lerp3(a, b, ti, v);
			memcpy(self + T_POS, v, 12);
		}
		if (flags[i] & 2)
		{
			lerp3(a + 3, b + 3, ti, v);
			memcpy(self + T_SCALE, v, 12);
		}
		if (flags[i] & 4)
		{
			const float* qa = a + 6; const float* qb = b + 6;
			float c[4] = { qb[0], qb[1], qb[2], qb[3] };
			float cosTheta = (qa[0] * qb[0] + qa[1] * qb[1]) + (qa[2] * qb[2] + qa[3] * qb[3]);
			if (cosTheta < 0.0f)
			{
				for (int l = 0; l < 4; l++) c[l] = -qb[l];
				cosTheta = -cosTheta;
			}
			if (cosTheta > 1.0f - 1.1920928955078125e-07f)
			{
				const float s = 1.0f - ti;
				for (int l = 0; l < 4; l++) v[l] = qa[l] * s + c[l] * ti;
			}
			else
			{
				const float angle = acosf(cosTheta);
				const float s0 = sinf((1.0f - ti) * angle), s1 = sinf(ti * angle), s2 = sinf(angle);
				for (int l = 0; l < 4; l++) v[l] = (qa[l] * s0 + c[l] * s1) / s2;
			}
			memcpy(self + T_ROT, v, 16);
		}
		if (flags[i] & 8)
		{
			const int active = roundf(ti) != 0.0f ? ((flags[i] >> 5) & 1) : ((flags[i] >> 4) & 1);
			if (oracle_set_active(transforms, stride, occupancy, &entityIds[i], 1, active) != 0)
				rc = -1;
		}
	}
	return rc;
}
