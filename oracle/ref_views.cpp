// TEST INFRASTRUCTURE ONLY — not part of the shipped product path.
//
// Second harness around UNMODIFIED reference code, for the view-setup rows:
//   calcLightViewProj is a file-static function of source/system/render/csm.cpp, so that translation unit is #included
//   where it lies under /root/reference and the function is called directly. Everything else csm.cpp defines (Vulkan
//   passes) is hidden and discarded at link time (-fvisibility=hidden -ffunction-sections -Wl,--gc-sections, oracle/Makefile),
//   which is why this library loads without the renderer.
//   The camera matrices of GraphicsSystem::prepareCommonConstants (source/system/graphics.cpp:168-172,192-203,241) cannot
//   be reached the same way (graphics.cpp needs the Vulkan headers); ref_camera_view_proj evaluates the SAME expressions
//   through the reference's own math headers.
// Nothing here is copied from the reference: this file only calls it.
#include "source/system/render/csm.cpp"

#define REF_EXPORT extern "C" __attribute__((visibility("default")))

REF_EXPORT void ref_light_view_proj(const float* view, const float* lightDir, float fov, float aspect, float nearPlane,
	float farPlane, float zCoeff, uint32_t shadowMapSize, float* viewProjOut, float* cameraOffsetOut)
{
	f32x4x4 v; memcpy(&v, view, 64);
	f32x4 offset = f32x4::zero;
	auto viewProj = calcLightViewProj(v, f32x4(lightDir[0], lightDir[1], lightDir[2], 0.0f), offset, fov, aspect, nearPlane,
		farPlane, zCoeff, shadowMapSize);
	memcpy(viewProjOut, &viewProj, 64); memcpy(cameraOffsetOut, &offset, 16);
}

// calcView of a camera without a parent (graphics.cpp:168-172), the camera-relative translation (:201), the projection of a
// perspective CameraComponent (camera.hpp:111-121) and viewProj = projection * view (:241).
REF_EXPORT void ref_camera_view_proj(const float* position, const float* rotation, const float* scaling, float fov, float aspect,
	float nearPlane, float* viewOut, float* projectionOut, float* viewProjOut)
{
	f32x4 p(position[0], position[1], position[2], 0.0f), s(scaling[0], scaling[1], scaling[2], 0.0f);
	quat q(rotation[0], rotation[1], rotation[2], rotation[3]);
	auto view = rotate(normalize(q)) * translate(scale(s), -p);
	setTranslation(view, f32x4::zero);
	auto projection = (f32x4x4)calcPerspProjInfRevZ(fov, aspect, nearPlane);
	auto viewProj = projection * view;
	memcpy(viewOut, &view, 64); memcpy(projectionOut, &projection, 64); memcpy(viewProjOut, &viewProj, 64);
}

// calcRelativeView (graphics.cpp:173-189) for a camera with `parentCount` ancestors, nearest first: the reference's own
// calcModel / operator* on each ancestor's (position, rotation, scale). Lane W of the position is ignored by translate(t);
// lane W of the scale is 0 here (a component holds childCapacity bits there), so `scale == f32x4::one` is never taken.
REF_EXPORT void ref_camera_view_proj_chain(const float* position, const float* rotation, const float* scaling, const float* parents,
	uint32_t parentCount, float fov, float aspect, float nearPlane, float* viewOut, float* projectionOut, float* viewProjOut)
{
	f32x4 p(position[0], position[1], position[2], 0.0f), s(scaling[0], scaling[1], scaling[2], 0.0f);
	quat q(rotation[0], rotation[1], rotation[2], rotation[3]);
	auto view = rotate(normalize(q)) * translate(scale(s), -p);
	for (uint32_t i = 0; i < parentCount; i++)
	{
		const float* a = parents + (size_t)i * 10;
		auto parentModel = calcModel(f32x4(a[0], a[1], a[2], 0.0f), quat(a[3], a[4], a[5], a[6]), f32x4(a[7], a[8], a[9], 0.0f));
		view = parentModel * view;
	}
	setTranslation(view, f32x4::zero);
	auto projection = (f32x4x4)calcPerspProjInfRevZ(fov, aspect, nearPlane);
	auto viewProj = projection * view;
	memcpy(viewOut, &view, 64); memcpy(projectionOut, &projection, 64); memcpy(viewProjOut, &viewProj, 64);
}

// The same with the projection of an orthographic CameraComponent (camera.hpp:119-120).
REF_EXPORT void ref_camera_view_proj_ortho(const float* position, const float* rotation, const float* scaling, const float* parents,
	uint32_t parentCount, const float* width, const float* height, const float* depth, float* viewOut, float* projectionOut,
	float* viewProjOut)
{
	f32x4 p(position[0], position[1], position[2], 0.0f), s(scaling[0], scaling[1], scaling[2], 0.0f);
	quat q(rotation[0], rotation[1], rotation[2], rotation[3]);
	auto view = rotate(normalize(q)) * translate(scale(s), -p);
	for (uint32_t i = 0; i < parentCount; i++)
	{
		const float* a = parents + (size_t)i * 10;
		auto parentModel = calcModel(f32x4(a[0], a[1], a[2], 0.0f), quat(a[3], a[4], a[5], a[6]), f32x4(a[7], a[8], a[9], 0.0f));
		view = parentModel * view;
	}
	setTranslation(view, f32x4::zero);
	auto projection = (f32x4x4)calcOrthoProjRevZ(float2(width[0], width[1]), float2(height[0], height[1]), float2(depth[0], depth[1]));
	auto viewProj = projection * view;
	memcpy(viewOut, &view, 64); memcpy(projectionOut, &projection, 64); memcpy(viewProjOut, &viewProj, 64);
}
