// TEST INFRASTRUCTURE ONLY — not part of the shipped product path.
//
// Headless harness around the UNMODIFIED reference translation units of cfnptr/garden
// (compiled where they lie under /root/reference by oracle/Makefile, output oracle/_ref/).
// It builds synthetic scenes through the reference's real ECS API and calls the reference's
// own MeshRenderSystem::prepareSystems()/prepareMeshes() (source/system/render/mesh.cpp:69-108,331-553),
// then exposes the private draw lists through a small C ABI so python/ctypes tests can compare
// them against the CUDA path and against the plain-C restatement (oracle/sceneprep_oracle.c).
//
// Compiled with -fno-access-control so that the private members named in SURVEY.md §8c are reachable.
// Three functions the reference TUs link against but never reach on this path are stubbed at the bottom.
//
// Nothing here is copied from the reference: this file only *calls* its public/private API.

#include "garden/system/render/mesh.hpp"
#include "garden/system/transform.hpp"
#include "garden/system/thread.hpp"
#include "garden/system/graphics.hpp"
#include "garden/system/log.hpp"
#include "math/frustum.hpp"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace garden;
using namespace ecsm;
using namespace math;

namespace
{

// A trivial concrete mesh system (the open-source tree ships none for 3-D meshes).
// K selects a distinct C++ type (the ECS keys systems by typeid); stride = 48 + 16 * (K % 3) bytes,
// so K%3 != 0 exercises getMeshComponentSize() != sizeof(MeshRenderComponent).
template<int K>
struct HarnessMeshComponent final : public MeshRenderComponent
{
	uint8 extra[(K % 3) == 0 ? 1 : 16 * (K % 3)] = {};
};
template<> struct HarnessMeshComponent<0> final : public MeshRenderComponent { };
template<> struct HarnessMeshComponent<3> final : public MeshRenderComponent { };

struct IHarnessPool
{
	MeshRenderType renderType = MeshRenderType::Opaque;
	bool drawReady = true;       // isDrawReady(shadowPass < 0): the base pipeline (instance.cpp:61-87)
	bool drawReadyShadow = true; // isDrawReady(shadowPass >= 0): the shadow pipeline (none by default; UI: label.cpp:262-265)
	std::vector<uint8> readyCounts; // per slot; empty = default behaviour (mesh.hpp:142-146)
	virtual ~IHarnessPool() { }
	virtual void* data() = 0;
	virtual uint32 stride() = 0;
	virtual uint32 occupancy() = 0;
	virtual uint32 count() = 0;
	virtual MeshRenderComponent* addTo(ID<Entity> entity) = 0;
};

template<int K>
class HarnessMeshSystem final : public ComponentSystem<HarnessMeshComponent<K>, false>,
	public IMeshRenderSystem, public IHarnessPool
{
public:
	using C = HarnessMeshComponent<K>;
	HarnessMeshSystem() { Manager::Instance::get()->addGroupSystem<IMeshRenderSystem>(this); }

	bool isDrawReady(int8 shadowPass) override { return shadowPass < 0 ? drawReady : drawReadyShadow; }
	void drawAsync(MeshRenderComponent*, const f32x4x4&, const f32x4x4&, uint32, int32) override { }
	MeshRenderType getMeshRenderType() const override { return renderType; }
	MeshRenderPool& getMeshComponentPool() const override { return *((MeshRenderPool*)&this->components); }
	psize getMeshComponentSize() const override { return sizeof(C); }

	// Mirrors what overriding systems do (e.g. source/system/render/sprite.cpp:90-97):
	// frustum test first, then a per-component resource-ready predicate.
	uint32 getReadyMeshesAsync(MeshRenderComponent* meshRenderView,
		const f32x4& cameraPosition, const Frustum& frustum, f32x4x4& model) override
	{
		if (readyCounts.empty())
			return IMeshRenderSystem::getReadyMeshesAsync(meshRenderView, cameraPosition, frustum, model);
		if (isBehindFrustum(frustum, meshRenderView->aabb, model))
			return 0;
		auto slot = (uint32)(((uint8*)meshRenderView - (uint8*)this->components.getData()) / sizeof(C));
		return slot < readyCounts.size() ? readyCounts[slot] : 1;
	}

	void* data() override { return (void*)this->components.getData(); }
	uint32 stride() override { return (uint32)sizeof(C); }
	uint32 occupancy() override { return this->components.getOccupancy(); }
	uint32 count() override { return this->components.getCount(); }
	MeshRenderComponent* addTo(ID<Entity> entity) override
	{
		auto view = Manager::Instance::get()->add<C>(entity);
		return (MeshRenderComponent*)*view;
	}
};

constexpr int maxPools = 6;

Manager* manager = nullptr;
GraphicsSystem* fakeGraphics = nullptr;
MeshRenderSystem* meshRender = nullptr;
TransformSystem* transformSystem = nullptr;
IHarnessPool* pools[maxPools] = {};
int poolCount = 0;
std::vector<ID<Entity>> entityIDs;
bool systemsPrepared = false;

template<int K> IHarnessPool* createPool()
{
	manager->createSystem<HarnessMeshSystem<K>>();
	return manager->get<HarnessMeshSystem<K>>();
}

} // namespace

extern "C"
{

// threads < 0: reference default (getBestForegroundThreadCount, include/garden/os.hpp:54-72);
// threads == 0: asyncPreparing = false (single-threaded path of prepareMeshes); threads > 0 is not
// selectable in the reference (the pool size is fixed by the CPU), so it is treated as "default".
int ref_init(int threads, int useOIT)
{
	if (manager)
		return -1;
	manager = new Manager();
	fakeGraphics = (GraphicsSystem*)calloc(1, sizeof(GraphicsSystem));
	Singleton<GraphicsSystem>::singletonInstance = fakeGraphics;
	if (threads != 0)
		manager->createSystem<ThreadSystem>();
	manager->createSystem<TransformSystem>();
	manager->createSystem<MeshRenderSystem>(useOIT != 0, false, threads != 0);
	transformSystem = manager->get<TransformSystem>();
	meshRender = manager->get<MeshRenderSystem>();
	poolCount = 0; systemsPrepared = false;
	entityIDs.clear();
	return 0;
}

int ref_thread_count()
{
	auto threadSystem = ThreadSystem::Instance::tryGet();
	return threadSystem ? (int)threadSystem->getForegroundPool().getThreadCount() : 1;
}

void ref_shutdown()
{
	if (!manager)
		return;
	// Unlink the hierarchy by hand first: the manager's destructor clears entities before systems, and
	// TransformComponent::destroy() (source/system/transform.cpp:29-70) would look up already-cleared parents.
	{
		const auto& components = transformSystem->getComponents();
		auto data = (TransformComponent*)components.getData();
		auto occupancy = components.getOccupancy();
		for (uint32 i = 0; i < occupancy; i++)
		{
			data[i].parent = {};
			if (data[i].childs)
				free(data[i].childs);
			data[i].childs = nullptr;
			data[i].childCount() = 0; data[i].childCapacity() = 0;
		}
	}
	delete manager; manager = nullptr;
	Singleton<GraphicsSystem>::singletonInstance = nullptr;
	free(fakeGraphics); fakeGraphics = nullptr;
	meshRender = nullptr; transformSystem = nullptr;
	for (auto& p : pools) p = nullptr;
	poolCount = 0; entityIDs.clear();
}

// renderType: MeshRenderType numeric value (include/garden/system/render/mesh.hpp:30-40).
// Pool k gets component stride 48 + 16 * (k % 3).
int ref_add_pool(int renderType)
{
	if (poolCount >= maxPools)
		return -1;
	IHarnessPool* pool = nullptr;
	switch (poolCount)
	{
	case 0: pool = createPool<0>(); break;
	case 1: pool = createPool<1>(); break;
	case 2: pool = createPool<2>(); break;
	case 3: pool = createPool<3>(); break;
	case 4: pool = createPool<4>(); break;
	case 5: pool = createPool<5>(); break;
	}
	pool->renderType = (MeshRenderType)renderType;
	pools[poolCount] = pool;
	systemsPrepared = false;
	return poolCount++;
}

void ref_set_pool_draw_ready(int pool, int ready) { pools[pool]->drawReady = pools[pool]->drawReadyShadow = ready != 0; }
// readiness per kind of pass, the distinction the reference's systems make (instance.cpp:61-113, label.cpp:262-265)
void ref_set_pool_draw_ready2(int pool, int readyMain, int readyShadow)
{
	pools[pool]->drawReady = readyMain != 0; pools[pool]->drawReadyShadow = readyShadow != 0;
}

// Entities are created in order; entity i gets ID i + 1 + (entities created earlier).
// position/scale: [count][3], rotation: [count][4] xyzw, parent: [count] index into all entities created so far
// (or -1), flags bit0 = has TransformComponent, bit1 = modelWithAncestors.
// Returns index of the first created entity.
int ref_create_entities(uint32_t count, const float* position, const float* rotation,
	const float* scale, const int32_t* parent, const uint8_t* flags)
{
	auto first = (int)entityIDs.size();
	entityIDs.reserve(entityIDs.size() + count);
	for (uint32_t i = 0; i < count; i++)
	{
		auto entity = manager->createEntity();
		entityIDs.push_back(entity);
		if (flags && !(flags[i] & 1))
			continue;
		auto transformView = manager->add<TransformComponent>(entity);
		transformView->setPosition(float3(position[i * 3], position[i * 3 + 1], position[i * 3 + 2]));
		transformView->setRotation(quat(rotation[i * 4], rotation[i * 4 + 1], rotation[i * 4 + 2], rotation[i * 4 + 3]));
		transformView->setScale(float3(scale[i * 3], scale[i * 3 + 1], scale[i * 3 + 2]));
		if (flags)
			transformView->modelWithAncestors = (flags[i] & 2) != 0;
	}
	// Hierarchy through setParent only (tryAddChild quirk).
	for (uint32_t i = 0; i < count; i++)
	{
		if (!parent || parent[i] < 0 || (flags && !(flags[i] & 1)))
			continue;
		auto transformView = manager->get<TransformComponent>(entityIDs[first + i]);
		transformView->setParent(entityIDs[parent[i]]);
	}
	return first;
}

// Overwrites TRS of existing entities (animated subset; plain stores like setPosition/..., transform.hpp:74-104).
void ref_update_trs(uint32_t count, const uint32_t* entityIndex, const float* position,
	const float* rotation, const float* scale)
{
	for (uint32_t i = 0; i < count; i++)
	{
		auto transformView = manager->get<TransformComponent>(entityIDs[entityIndex[i]]);
		transformView->setPosition(float3(position[i * 3], position[i * 3 + 1], position[i * 3 + 2]));
		transformView->setRotation(quat(rotation[i * 4], rotation[i * 4 + 1], rotation[i * 4 + 2], rotation[i * 4 + 3]));
		transformView->setScale(float3(scale[i * 3], scale[i * 3 + 1], scale[i * 3 + 2]));
	}
}

void ref_set_active(uint32_t count, const uint32_t* entityIndex, int active)
{
	for (uint32_t i = 0; i < count; i++)
		manager->get<TransformComponent>(entityIDs[entityIndex[i]])->setActive(active != 0);
}

// aabb: [count][6] = min xyz, max xyz. enabled: nullable. readyCount: nullable (then default frustum-only predicate).
void ref_add_meshes(int pool, uint32_t count, const uint32_t* entityIndex, const float* aabb,
	const uint8_t* enabled, const uint8_t* readyCount)
{
	auto p = pools[pool];
	for (uint32_t i = 0; i < count; i++)
	{
		auto mesh = p->addTo(entityIDs[entityIndex[i]]);
		mesh->aabb = Aabb(f32x4(aabb[i * 6], aabb[i * 6 + 1], aabb[i * 6 + 2]),
			f32x4(aabb[i * 6 + 3], aabb[i * 6 + 4], aabb[i * 6 + 5]));
		if (enabled)
			mesh->isEnabled = enabled[i] != 0;
		if (readyCount)
		{
			auto slot = (uint32)(((uint8*)mesh - (uint8*)p->data()) / p->stride());
			if (p->readyCounts.size() <= slot)
				p->readyCounts.resize(slot + 1, 1);
			p->readyCounts[slot] = readyCount[i];
		}
	}
}

// Destroys whole entities (with all their components) and disposes, leaving freed pool slots (entity == 0).
void ref_destroy_entities(uint32_t count, const uint32_t* entityIndex)
{
	for (uint32_t i = 0; i < count; i++)
	{
		auto& entity = entityIDs[entityIndex[i]];
		// Detach from the hierarchy first, the way the engine does before destroying a node.
		auto transformView = manager->tryGet<TransformComponent>(entity);
		if (transformView)
			transformView->setParent({});
		manager->destroy(entity);
	}
	// same order as Manager::update (libraries/ecsm/source/ecsm.cpp:591-593); repeated until the pools are quiescent,
	// because disposing an entity only queues its components for the next system dispose
	for (int pass = 0; pass < 2; pass++)
	{
		manager->disposeGarbageComponents();
		manager->disposeEntities();
		manager->disposeSystemComponents();
	}
}

void ref_transform_pool(const void** data, uint32_t* stride, uint32_t* occupancy)
{
	const auto& components = transformSystem->getComponents();
	*data = components.getData(); *stride = (uint32_t)sizeof(TransformComponent);
	*occupancy = components.getOccupancy();
}
void ref_mesh_pool(int pool, void** data, uint32_t* stride, uint32_t* occupancy, uint32_t* count)
{
	*data = pools[pool]->data(); *stride = pools[pool]->stride();
	*occupancy = pools[pool]->occupancy(); *count = pools[pool]->count();
}
const uint8_t* ref_pool_ready_counts(int pool, uint32_t* size)
{
	*size = (uint32_t)pools[pool]->readyCounts.size();
	return pools[pool]->readyCounts.data();
}

void ref_set_camera(const float* cameraPos)
{
	fakeGraphics->commonConstants.cameraPos = float3(cameraPos[0], cameraPos[1], cameraPos[2]);
}

// planes out: [6][4] (normal xyz, distance), from the reference's Frustum(viewProj) (math/frustum.hpp:51-61).
// viewProj: 16 floats, column-major (c0..c3).
void ref_frustum_planes(const float* viewProj, float* planes)
{
	f32x4x4 m(f32x4(viewProj[0], viewProj[1], viewProj[2], viewProj[3]),
		f32x4(viewProj[4], viewProj[5], viewProj[6], viewProj[7]),
		f32x4(viewProj[8], viewProj[9], viewProj[10], viewProj[11]),
		f32x4(viewProj[12], viewProj[13], viewProj[14], viewProj[15]));
	Frustum frustum(m);
	static_assert(sizeof(Plane) == 16, "plane layout");
	memcpy(planes, frustum.planes, 6 * sizeof(Plane));
}

static Frustum frustumFromPlanes(const float* planes, int planeCount)
{
	Frustum frustum;
	memcpy(frustum.planes, planes, 6 * sizeof(Plane));
	frustum.count = (uint8)planeCount;
	return frustum;
}

// One reference prepareMeshes call (mesh.cpp:331-553). planes: [6][4]; uiPlanes nullable.
void ref_prepare(const float* planes, int planeCount, const float* uiPlanes, int uiPlaneCount,
	const float* cameraOffset, int shadowPass)
{
	if (!systemsPrepared)
	{
		meshRender->prepareSystems(); // mesh.cpp:69-108
		systemsPrepared = true;
	}
	auto viewFrustum = frustumFromPlanes(planes, planeCount);
	Frustum uiFrustum;
	if (uiPlanes)
		uiFrustum = frustumFromPlanes(uiPlanes, uiPlaneCount);
	meshRender->prepareMeshes(viewFrustum, uiPlanes ? &uiFrustum : nullptr,
		f32x4(cameraOffset[0], cameraOffset[1], cameraOffset[2], cameraOffset[3]), (int8)shadowPass);
}

uint32_t ref_unsorted_buffer_count() { return meshRender->unsortedBufferCount; }
uint32_t ref_sorted_buffer_count() { return meshRender->sortedBufferCount; }

// records: 64-byte UnsortedMesh (mesh.hpp:191-197)
void ref_get_unsorted(uint32_t buffer, const void** records, uint32_t* drawCount, uint32_t* instanceCount)
{
	static_assert(sizeof(MeshRenderSystem::UnsortedMesh) == 64, "record layout");
	static_assert(sizeof(MeshRenderSystem::SortedMesh) == 64, "record layout");
	auto unsortedBuffer = meshRender->unsortedBuffers[buffer];
	*records = unsortedBuffer->combinedMeshes.data();
	*drawCount = unsortedBuffer->drawCount.load();
	*instanceCount = unsortedBuffer->instanceCount.load();
}
void ref_get_sorted_counts(uint32_t buffer, uint32_t* drawCount, uint32_t* instanceCount)
{
	*drawCount = meshRender->sortedBuffers[buffer]->drawCount.load();
	*instanceCount = meshRender->sortedBuffers[buffer]->instanceCount.load();
}
// records: 64-byte SortedMesh (mesh.hpp:198-205); which: 0 = translucent list, 1 = UI list
void ref_get_sorted(int which, const void** records, uint32_t* drawCount)
{
	if (which == 0)
	{
		*records = meshRender->transSortedMeshes.data(); *drawCount = meshRender->transDrawIndex.load();
	}
	else
	{
		*records = meshRender->uiSortedMeshes.data(); *drawCount = meshRender->uiDrawIndex.load();
	}
}

// World matrix of one entity straight from TransformComponent::calcModel (transform.hpp:197-214); out: 16 floats.
void ref_calc_model(uint32_t entityIndex, const float* cameraPos, float* out)
{
	auto transformView = manager->get<TransformComponent>(entityIDs[entityIndex]);
	auto model = transformView->calcModel(f32x4(cameraPos[0], cameraPos[1], cameraPos[2], 0.0f));
	memcpy(out, &model, 64);
}

// Instance data of a draw list the way renderUnsorted / renderSorted + drawAsync produce it (mesh.cpp:600-603,
// sprite.cpp:122-130): model = f32x4x4(bakedModel, (0,0,0,1)); mvp = (float4x4)(viewProj * model). out: count x 16 floats.
void ref_instance_mvp(const float* viewProj, const void* records, uint32_t count, float* out)
{
	f32x4x4 vp;
	memcpy(&vp, viewProj, 64);
	auto meshes = (const MeshRenderSystem::UnsortedMesh*)records;
	for (uint32_t i = 0; i < count; i++)
	{
		auto model = f32x4x4(meshes[i].bakedModel, f32x4(0.0f, 0.0f, 0.0f, 1.0f));
		auto mvp = (float4x4)(vp * model);
		memcpy(out + (size_t)i * 16, &mvp, 64);
	}
}

// f2: TransformSystem::animateAsync (source/system/transform.cpp:609-623) for `count` entities, through the real system and
// the real TransformFrame type. flags[i]: bit0 animatePosition, bit1 animateScale, bit2 animateRotation, bit3 animateIsActive,
// bit4 frameA.isActive, bit5 frameB.isActive. frameA / frameB: [count][10] = position xyz, scale xyz, rotation xyzw.
void ref_animate(uint32_t count, const uint32_t* entityIndex, const uint8_t* flags, const float* frameA, const float* frameB,
	const float* t)
{
	for (uint32_t i = 0; i < count; i++)
	{
		TransformFrame a, b;
		const float* fa = frameA + (size_t)i * 10; const float* fb = frameB + (size_t)i * 10;
		a.animatePosition = b.animatePosition = (flags[i] >> 0) & 1;
		a.animateScale = b.animateScale = (flags[i] >> 1) & 1;
		a.animateRotation = b.animateRotation = (flags[i] >> 2) & 1;
		a.animateIsActive = b.animateIsActive = (flags[i] >> 3) & 1;
		a.isActive = (flags[i] >> 4) & 1; b.isActive = (flags[i] >> 5) & 1;
		a.position = f32x4(fa[0], fa[1], fa[2], 0.0f); b.position = f32x4(fb[0], fb[1], fb[2], 0.0f);
		a.scale = f32x4(fa[3], fa[4], fa[5], 0.0f); b.scale = f32x4(fb[3], fb[4], fb[5], 0.0f);
		a.rotation = quat(fa[6], fa[7], fa[8], fa[9]); b.rotation = quat(fb[6], fb[7], fb[8], fb[9]);
		auto transformView = manager->get<TransformComponent>(entityIDs[entityIndex[i]]);
		transformSystem->animateAsync(View<Component>(transformView), View<AnimationFrame>(&a), View<AnimationFrame>(&b), t[i]);
	}
}

// Times `frames` frames; one frame = viewCount serial prepareMeshes calls (mesh.cpp:795-847,893-903 order:
// shadow passes first, main view last is the caller's choice of ordering in the arrays).
// planes: [viewCount][6][4], cameraOffsets: [viewCount][4], shadowPasses: [viewCount]. Writes per-frame ms.
void ref_time_frames(uint32_t viewCount, const float* planes, const uint8_t* planeCounts,
	const float* cameraOffsets, const int8_t* shadowPasses, uint32_t frames, double* msOut, uint64_t* visibleOut)
{
	if (!systemsPrepared)
	{
		meshRender->prepareSystems();
		systemsPrepared = true;
	}
	for (uint32_t f = 0; f < frames; f++)
	{
		uint64_t visible = 0;
		auto t0 = std::chrono::steady_clock::now();
		for (uint32_t v = 0; v < viewCount; v++)
		{
			auto frustum = frustumFromPlanes(planes + v * 24, planeCounts[v]);
			meshRender->prepareMeshes(frustum, nullptr, f32x4(cameraOffsets[v * 4], cameraOffsets[v * 4 + 1],
				cameraOffsets[v * 4 + 2], cameraOffsets[v * 4 + 3]), shadowPasses[v]);
			for (uint32_t i = 0; i < meshRender->unsortedBufferCount; i++)
				visible += meshRender->unsortedBuffers[i]->drawCount.load();
			visible += meshRender->transDrawIndex.load();
		}
		auto t1 = std::chrono::steady_clock::now();
		msOut[f] = std::chrono::duration<double, std::milli>(t1 - t0).count();
		if (visibleOut)
			visibleOut[f] = visible;
	}
}

} // extern "C"

//**********************************************************************************************************************
// Stubs for symbols referenced by mesh.cpp / thread.cpp but unreachable on this path.
void GraphicsSystem::startRecording(CommandBufferType commandBufferType) noexcept { abort(); }
void GraphicsSystem::stopRecording() noexcept { abort(); }
void LogSystem::log(LogLevel level, string_view message) noexcept { }
