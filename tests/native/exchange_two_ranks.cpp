// Two GPUs, one process, no Python anywhere: drives the multi-GPU path of libgarden_sceneprep.so purely through the C ABI
// (include/garden_sceneprep.h). One thread per GPU (NCCL's rule for several devices in one process): each prepares a
// contiguous entity range and exchanges (gsp_comm_init_all, gsp_exchange_async); a third context then sorts the WHOLE scene
// on GPU 0. The concatenated merged slices must equal that single-GPU sort: key bits and slot order, every list.
//
//   g++ -std=c++17 -O2 -I include tests/native/exchange_two_ranks.cpp -L garden_b200 -lgarden_sceneprep -lpthread -o build/exchange_two_ranks
//   LD_LIBRARY_PATH=garden_b200 ./build/exchange_two_ranks [ranks]
#include "garden_sceneprep.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CHECK(call) do { int rc__ = (call); if (rc__ != GSP_OK) { fprintf(stderr, "%s failed with %d: %s\n", #call, rc__, gsp_last_error(ctx)); exit(2); } } while (0)

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
static float unit(uint32_t& s) { return (float)lcg(s) * (1.0f / 16777216.0f); }

struct Scene
{
	std::vector<uint8_t> transforms, meshes; // TransformComponent (80 B) / MeshRenderComponent (48 B) bytes
	uint32_t count = 0;
};
// entities [first, first + count) of the global scene; chains of 5 (every fifth entity a root), ids local to the shard
static Scene makeShard(uint32_t first, uint32_t count)
{
	Scene sc; sc.count = count;
	sc.transforms.assign((size_t)count * 80, 0); sc.meshes.assign((size_t)count * 48, 0);
	for (uint32_t i = 0; i < count; i++)
	{
		uint32_t seed = (first + i) * 2654435761u + 12345u;
		uint8_t* t = &sc.transforms[(size_t)i * 80];
		uint8_t* m = &sc.meshes[(size_t)i * 48];
		const uint32_t entity = i + 1, parent = ((first + i) % 5) ? i : 0; // parent id = previous entity (local, 1-based)
		memcpy(t + 0, &entity, 4); memcpy(t + 4, &parent, 4);
		float pos[3], scl[3], rot[4];
		const bool root = parent == 0;
		pos[0] = root ? (unit(seed) - 0.5f) * 400.0f : (unit(seed) - 0.5f) * 4.0f;
		pos[1] = root ? (unit(seed) - 0.5f) * 10.0f : (unit(seed) - 0.5f) * 4.0f;
		pos[2] = root ? (unit(seed) - 0.5f) * 400.0f : (unit(seed) - 0.5f) * 4.0f;
		for (float& v : scl) v = 0.6f + 0.8f * unit(seed);
		float len = 0.0f;
		for (float& v : rot) { v = unit(seed) * 2.0f - 1.0f; len += v * v; }
		rot[3] += 0.25f;
		memcpy(t + 16, pos, 12); memcpy(t + 32, scl, 12); memcpy(t + 48, rot, 16);
		t[72] = 1; t[73] = 1; t[74] = 1; // selfActive, ancestorsActive, modelWithAncestors
		memcpy(m + 0, &entity, 4);
		m[14] = 1; // isEnabled
		const float mn[4] = { -0.5f, -0.5f, -0.5f, 0.0f }, mx[4] = { 0.5f, 0.5f, 0.5f, 0.0f };
		memcpy(m + 16, mn, 16); memcpy(m + 32, mx, 16);
	}
	return sc;
}

static void makeViews(gsp_view views[2])
{
	// infinite reversed-Z perspective (libraries/math/include/math/matrix/projection.hpp:39-56), camera-relative view = identity
	const float t = std::tan(0.6f), aspect = 16.0f / 9.0f, nearPlane = 0.01f;
	float vp[16] = {};
	vp[0] = 1.0f / (aspect * t); vp[5] = -1.0f / t; vp[11] = 1.0f; vp[14] = nearPlane; // column-major
	const float zero[4] = { 0, 0, 0, 0 }, off[4] = { 3.0f, -1.0f, 2.0f, 0.0f };
	gsp_view_from_viewproj(vp, zero, -1, &views[1]);
	// a second view (a "shadow pass" with a camera offset): the same projection looking along +x, vq = vp * R with
	// R e_x = (0,0,1), R e_y = (0,1,0), R e_z = (-1,0,0)  =>  columns: vq.c0 = vp.c2, vq.c1 = vp.c1, vq.c2 = -vp.c0, vq.c3 = vp.c3
	float vq[16] = {};
	for (int i = 0; i < 4; i++)
	{
		vq[0 + i] = vp[8 + i]; vq[4 + i] = vp[4 + i]; vq[8 + i] = -vp[0 + i]; vq[12 + i] = vp[12 + i];
	}
	gsp_view_from_viewproj(vq, off, 0, &views[0]);
}

struct Slices { std::vector<uint32_t> keys, pays; std::vector<uint8_t> ranks; uint32_t start = 0; };

int main(int argc, char** argv)
{
	const uint32_t ranks = argc > 1 ? (uint32_t)atoi(argv[1]) : 2u, perRank = 100000; // (a multiple of the chain length)
	gsp_view views[2];
	makeViews(views);
	const float cam[3] = { 1.5f, 0.25f, -2.0f };
	std::vector<gsp_context*> ctxs(ranks, nullptr);
	for (uint32_t r = 0; r < ranks; r++)
	{
		gsp_context* ctx = nullptr;
		if (gsp_create((int)r, &ctxs[r]) != GSP_OK) { fprintf(stderr, "gsp_create(%u): %s\n", r, gsp_last_error(nullptr)); return 3; }
		ctx = ctxs[r];
		Scene sc = makeShard(r * perRank, perRank);
		CHECK(gsp_set_transforms(ctx, sc.transforms.data(), 80, sc.count));
		CHECK(gsp_set_pool_count(ctx, 1));
		CHECK(gsp_set_mesh_pool(ctx, 0, GSP_RT_OPAQUE, 1, sc.meshes.data(), 48, sc.count, sc.count, nullptr));
		CHECK(gsp_set_views(ctx, 2, views, cam));
	}
	{
		gsp_context* ctx = ctxs[0];
		CHECK(gsp_comm_init_all(ctxs.data(), ranks));
	}
	const uint32_t lists = 2;
	std::vector<std::vector<Slices>> got(ranks, std::vector<Slices>(lists));
	std::vector<std::thread> threads;
	for (uint32_t r = 0; r < ranks; r++)
		threads.emplace_back([&, r] {
			gsp_context* ctx = ctxs[r];
			CHECK(gsp_exchange_autosize(ctx, nullptr)); // collective
			for (int frame = 0; frame < 3; frame++)
			{
				CHECK(gsp_run_async(ctx));
				CHECK(gsp_exchange_async(ctx));
			}
			uint32_t bits = 0, need = 0;
			CHECK(gsp_exchange_finish(ctx, &bits, &need));
			if (bits) { fprintf(stderr, "rank %u: exchange flags %u (needed %u)\n", r, bits, need); exit(4); }
			CHECK(gsp_sync(ctx));
			for (uint32_t l = 0; l < lists; l++)
			{
				const uint32_t *k, *p; const uint8_t* rk; uint32_t start, count;
				CHECK(gsp_get_merged_device(ctx, l, &k, &p, &rk, &start, &count));
				Slices& s = got[r][l];
				s.start = start; s.keys.resize(count); s.pays.resize(count); s.ranks.resize(count);
				CHECK(gsp_copy_to_host(ctx, k, s.keys.data(), (size_t)count * 4));
				CHECK(gsp_copy_to_host(ctx, p, s.pays.data(), (size_t)count * 4));
				CHECK(gsp_copy_to_host(ctx, rk, s.ranks.data(), count));
			}
		});
	for (auto& t : threads) t.join();

	// the whole scene on one GPU
	gsp_context* ctx = nullptr;
	if (gsp_create(0, &ctx) != GSP_OK) return 3;
	Scene whole = makeShard(0, perRank * ranks);
	CHECK(gsp_set_transforms(ctx, whole.transforms.data(), 80, whole.count));
	CHECK(gsp_set_pool_count(ctx, 1));
	CHECK(gsp_set_mesh_pool(ctx, 0, GSP_RT_OPAQUE, 1, whole.meshes.data(), 48, whole.count, whole.count, nullptr));
	CHECK(gsp_set_views(ctx, 2, views, cam));
	CHECK(gsp_run(ctx));
	uint64_t checked = 0;
	for (uint32_t l = 0; l < lists; l++)
	{
		const uint32_t *dk, *dp; uint32_t count;
		CHECK(gsp_get_sorted_run_device(ctx, l, 0, 0, &dk, &dp, &count));
		std::vector<uint32_t> wk(count), wp(count);
		CHECK(gsp_copy_to_host(ctx, dk, wk.data(), (size_t)count * 4));
		CHECK(gsp_copy_to_host(ctx, dp, wp.data(), (size_t)count * 4));
		uint32_t pos = 0;
		for (uint32_t r = 0; r < ranks; r++)
		{
			const Slices& s = got[r][l];
			if (s.start != pos) { fprintf(stderr, "list %u: rank %u slice starts at %u, expected %u\n", l, r, s.start, pos); return 5; }
			for (size_t i = 0; i < s.keys.size(); i++, pos++)
			{
				const uint32_t slot = (s.pays[i] & 0x0FFFFFFFu) + s.ranks[i] * perRank; // shards are contiguous ranges
				if (pos >= count || s.keys[i] != wk[pos] || slot != (wp[pos] & 0x0FFFFFFFu))
				{
					fprintf(stderr, "list %u position %u: merged (key %08x, slot %u) vs single GPU (key %08x, slot %u)\n", l, pos,
						s.keys[i], slot, pos < count ? wk[pos] : 0u, pos < count ? wp[pos] & 0x0FFFFFFFu : 0u);
					return 6;
				}
			}
		}
		if (pos != count) { fprintf(stderr, "list %u: slices cover %u of %u\n", l, pos, count); return 7; }
		checked += count;
	}
	uint32_t a2a = 0;
	gsp_comm_info(ctxs[0], nullptr, nullptr, nullptr, &a2a);
	printf("exchange_two_ranks ok: %u ranks x %u entities, %u lists, %llu merged elements == single-GPU sort (%s, NCCL inside the library, no Python)\n",
		ranks, perRank, lists, (unsigned long long)checked, a2a ? "alltoall" : "allgather");
	if (checked < 1000) { fprintf(stderr, "suspiciously few visible elements\n"); return 8; }
	gsp_destroy(ctx);
	for (auto c : ctxs) gsp_destroy(c);
	return 0;
}
