/* TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference's scene-preparation path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (garden_b200/) never links, imports or calls anything under oracle/.
 *
 * Parity status: PINNED — this restatement is checked bit-for-bit against the reference's own translation units
 * (oracle/_ref/libgarden_ref_parity.so, built by oracle/Makefile from /root/reference) by tests/test_oracle_vs_ref.py,
 * and against golden vectors generated from that build (tests/golden/, tests/golden/make_golden.py).
 * The reference's own tests hold no golden vectors for this path.
 */
#ifndef SCENEPREP_ORACLE_H
#define SCENEPREP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_POOLS 8

/* MeshRenderType, include/garden/system/render/mesh.hpp:30-40 */
enum { ORACLE_RT_COLOR = 0, ORACLE_RT_OPAQUE, ORACLE_RT_TRANSLUCENT, ORACLE_RT_OIT, ORACLE_RT_REFRACTED,
	ORACLE_RT_TRANS_DEPTH, ORACLE_RT_UI, ORACLE_RT_COUNT };

/* UnsortedMesh / SortedMesh, include/garden/system/render/mesh.hpp:191-205 (64 bytes each).
 * bufferIndex is padding (left 0) in UnsortedMesh. */
typedef struct OracleRecord
{
	uint64_t componentOffset;
	float bakedModel[12]; /* float4x3: columns c0..c3, lanes xyz */
	float distanceSq;
	uint32_t bufferIndex;
} OracleRecord;

typedef struct OracleView
{
	float planes[6][4];   /* normal xyz, distance; unnormalised (math/frustum.hpp:51-61) */
	uint32_t planeCount;
	float uiPlanes[6][4];
	uint32_t uiPlaneCount; /* 0 = no UI frustum given */
	float cameraOffset[4];
	int32_t shadowPass;    /* < 0 = main view */
} OracleView;

typedef struct OracleScene OracleScene;

OracleScene* oracle_create(void);
void oracle_destroy(OracleScene* s);

/* Pointers are borrowed (no copy); layouts are the reference's AoS pools. */
int oracle_set_transforms(OracleScene* s, const void* data, uint32_t stride, uint32_t occupancy);
int oracle_set_pool(OracleScene* s, uint32_t index, uint32_t renderType, uint32_t drawReady, void* data,
	uint32_t stride, uint32_t occupancy, uint32_t count, const uint8_t* readyCounts, uint32_t readyCountsSize);
void oracle_set_pool_count(OracleScene* s, uint32_t poolCount);
/* IMeshRenderSystem::isDrawReady(shadowPass) is asked once per prepareMeshes call (mesh.cpp:426,482) and systems answer
 * per kind of pass (instance.cpp:61-113: base pipeline for shadowPass < 0, shadow pipeline otherwise; label.cpp:262-265):
 * readiness oracle_prepare uses for main views / for shadow passes. oracle_set_pool sets both to its drawReady argument. */
void oracle_set_pool_draw_ready(OracleScene* s, uint32_t pool, uint32_t readyMain, uint32_t readyShadow);
void oracle_set_camera(OracleScene* s, const float cameraPos[3]);

/* One MeshRenderSystem::prepareMeshes call (mesh.cpp:331-553) including sortMeshes (mesh.cpp:265-328).
 * writeVisible != 0 stores MeshRenderComponent::isVisible into the AoS pool like the reference does. */
int oracle_prepare(OracleScene* s, const OracleView* view, int writeVisible);

uint32_t oracle_unsorted_buffer_count(const OracleScene* s);
uint32_t oracle_sorted_buffer_count(const OracleScene* s);
void oracle_get_unsorted(const OracleScene* s, uint32_t buffer, const OracleRecord** records,
	uint32_t* drawCount, uint32_t* instanceCount);
void oracle_get_sorted_counts(const OracleScene* s, uint32_t buffer, uint32_t* drawCount, uint32_t* instanceCount);
void oracle_get_sorted(const OracleScene* s, int which, const OracleRecord** records, uint32_t* drawCount);
/* Per-slot visibility of the last main-view prepare (0/1), 0xFF where the reference would not have written. */
const uint8_t* oracle_get_visible(const OracleScene* s, uint32_t pool);

/* TransformComponent::calcModel(cameraPosition) for one transform slot (transform.hpp:197-214); out: 16 floats. */
int oracle_calc_model(const OracleScene* s, uint32_t transformSlot, const float cameraPos[3], float out[16]);
/* Frustum(viewProj) (math/frustum.hpp:51-61); viewProj column-major 16 floats; planes out [6][4]. */
void oracle_frustum_planes(const float viewProj[16], float planes[6][4]);
/* f32x4x4 * f32x4x4 (simd/matrix/float.hpp:193-220), column-major. */
void oracle_mat_mul(const float a[16], const float b[16], float out[16]);
/* math::calcModel(position, rotation, scale) general branch (matrix/transform.hpp:251-256). */
void oracle_local_model(const float pos[3], const float rot[4], const float scale[3], float out[16]);

/* ---- SURVEY.md §8f rows ("next") ------------------------------------------------------------------------------------ */
/* f1: per-instance mvp = (float4x4)(viewProj * f32x4x4(bakedModel, (0,0,0,1))) for every record of a draw list, in draw
 * order (mesh.cpp:600-603 + sprite.cpp:122-130). viewProj column-major; out: count x 16 floats. */
void oracle_instance_mvp(const float viewProj[16], const OracleRecord* records, uint32_t count, float* out);
/* f3: TransformComponent::setActive (source/system/transform.cpp:75-127) applied to `count` entities in order, on the AoS
 * transform pool in place (selfActive @72, ancestorsActive @73): explicit stack walk over child lists rebuilt from the
 * parent links, exactly as the reference walks `childs`. entityIds are 1-based ECS ids. Returns 0, or -1 on bad input. */
int oracle_set_active(void* transforms, uint32_t stride, uint32_t occupancy, const uint32_t* entityIds, uint32_t count,
	int active);

/* f2 (oracle only so far — no CUDA counterpart yet): TransformSystem::animateAsync (source/system/transform.cpp:609-623) for
 * `count` entities on the AoS transform pool in place: position / scale = lerp(a, b, t) (simd/vector/float.hpp:1469),
 * rotation = slerp(a, b, t) (quaternion.hpp:175-193; acosf / sinf come from the host libm exactly as in the reference),
 * isActive = round(t) ? b : a through setActive. flags[i]: bit0 position, bit1 scale, bit2 rotation, bit3 isActive animated,
 * bit4 frameA.isActive, bit5 frameB.isActive; frameA / frameB: [count][10] = position xyz, scale xyz, rotation xyzw. */
int oracle_animate(void* transforms, uint32_t stride, uint32_t occupancy, uint32_t count, const uint32_t* entityIds,
	const uint8_t* flags, const float* frameA, const float* frameB, const float* t);

#ifdef __cplusplus
}
#endif
#endif
