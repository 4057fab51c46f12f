// View setup on the host: the matrices and frusta prepareMeshes receives, restated in the reference's
// operation order so that the planes — and with them every visibility decision — are bit-identical to the reference's:
//   camera      GraphicsSystem::prepareCommonConstants   source/system/graphics.cpp:168-172,192-203,241  (view, projection, viewProj)
//   cascades    calcLightViewProj + prepareShadowRender   source/system/render/csm.cpp:260-308,311-329
// and what they call in libraries/math: calcPerspProj[Inf]RevZ / calcOrthoProjRevZ (matrix/projection.hpp:39-99), lookAt
// (matrix/transform.hpp:291-299), rotate / translate / scale (matrix/transform.hpp:50-140), f32x4x4 products and inverse4x4
// (simd/matrix/float.hpp:193-242,472-544), dot3 / cross3 / normalize3 / normalize4 (simd/vector/float.hpp:1090-1215).
// Everything is float arithmetic of the reference's x86 build without contraction ("dialect B", SURVEY.md finding 3): four
// lanes at a time, an FMA exactly where the reference's AVX2 build fuses (MATH_SIMD_FMA), dpps sums as (p0 + p1) + (p2 + p3),
// tan / floor from the host libm like the reference. Host code only: no device is needed for these entry points.
// Pinned against the reference's OWN csm.cpp and math headers (oracle/ref_views.cpp -> oracle/_ref/libgarden_ref_views.so,
// tests/test_views.py) and against golden vectors made with them (tests/golden/frows/views.npz).
#include "../../include/garden_sceneprep.h"

#include <cmath>
#include <cstring>

namespace
{

struct L4 { float v[4]; };          // one SSE register
struct M4 { L4 c[4]; };             // f32x4x4: columns c0..c3

// (volatile stores keep every operation a separately rounded float operation whatever the host compiler flags are)
inline float rnd(float x) { volatile float y = x; return y; }
inline L4 make(float x, float y, float z, float w) { return L4{{x, y, z, w}}; }
inline L4 splat(float x) { return make(x, x, x, x); }
inline L4 mul(const L4& a, const L4& b) { L4 r; for (int i = 0; i < 4; i++) r.v[i] = rnd(a.v[i] * b.v[i]); return r; }
inline L4 add(const L4& a, const L4& b) { L4 r; for (int i = 0; i < 4; i++) r.v[i] = rnd(a.v[i] + b.v[i]); return r; }
inline L4 sub(const L4& a, const L4& b) { L4 r; for (int i = 0; i < 4; i++) r.v[i] = rnd(a.v[i] - b.v[i]); return r; }
inline L4 dvd(const L4& a, const L4& b) { L4 r; for (int i = 0; i < 4; i++) r.v[i] = rnd(a.v[i] / b.v[i]); return r; }
inline L4 fma4(const L4& a, const L4& b, const L4& c) { L4 r; for (int i = 0; i < 4; i++) r.v[i] = std::fmaf(a.v[i], b.v[i], c.v[i]); return r; }
inline L4 neg(const L4& a) { L4 r; for (int i = 0; i < 4; i++) r.v[i] = -a.v[i]; return r; }
// _mm_min_ps / _mm_max_ps: (a < b) ? a : b and (a > b) ? a : b, the second operand on ties / NaN
inline L4 min4(const L4& a, const L4& b) { L4 r; for (int i = 0; i < 4; i++) r.v[i] = a.v[i] < b.v[i] ? a.v[i] : b.v[i]; return r; }
inline L4 max4(const L4& a, const L4& b) { L4 r; for (int i = 0; i < 4; i++) r.v[i] = a.v[i] > b.v[i] ? a.v[i] : b.v[i]; return r; }
// _mm_shuffle_ps(a, b, _MM_SHUFFLE(i3, i2, i1, i0)) = (a[i0], a[i1], b[i2], b[i3])
inline L4 shuf(const L4& a, const L4& b, int i0, int i1, int i2, int i3) { return make(a.v[i0], a.v[i1], b.v[i2], b.v[i3]); }
// _mm_dp_ps(a, b, 0x7f) / (.., 0xff) in every lane: (p0 + p1) + (p2 + p3), p3 = +0 for the 3-lane form
inline float dp3(const L4& a, const L4& b)
{
	return rnd(rnd(rnd(a.v[0] * b.v[0]) + rnd(a.v[1] * b.v[1])) + rnd(rnd(a.v[2] * b.v[2]) + 0.0f));
}
inline float dp4(const L4& a, const L4& b)
{
	return rnd(rnd(rnd(a.v[0] * b.v[0]) + rnd(a.v[1] * b.v[1])) + rnd(rnd(a.v[2] * b.v[2]) + rnd(a.v[3] * b.v[3])));
}
inline L4 normalize3(const L4& v) { return dvd(v, splat(rnd(std::sqrt(dp3(v, v))))); }    // simd/vector/float.hpp:1212-1215
inline L4 normalize4(const L4& v) { return dvd(v, splat(rnd(std::sqrt(dp4(v, v))))); }    // :1198-1201
inline L4 cross3(const L4& a, const L4& b)                                                  // :1107-1113
{
	const L4 r = sub(mul(shuf(b, b, 1, 2, 0, 0), a), mul(shuf(a, a, 1, 2, 0, 0), b));
	return shuf(r, r, 1, 2, 0, 0);
}
// f32x4x4 * f32x4 and f32x4x4 * f32x4x4: mul, fma, fma, fma per column (simd/matrix/float.hpp:193-204,225-231)
inline L4 mulVec(const M4& m, const L4& v)
{
	L4 r = mul(m.c[0], splat(v.v[0]));
	r = fma4(m.c[1], splat(v.v[1]), r);
	r = fma4(m.c[2], splat(v.v[2]), r);
	return fma4(m.c[3], splat(v.v[3]), r);
}
inline M4 mulMat(const M4& a, const M4& b)
{
	M4 r;
	for (int i = 0; i < 4; i++) r.c[i] = mulVec(a, b.c[i]);
	return r;
}
inline M4 transpose(const M4& m)
{
	M4 r;
	for (int i = 0; i < 4; i++) for (int l = 0; l < 4; l++) r.c[i].v[l] = m.c[l].v[i];
	return r;
}
// float4x4(row-wise arguments) -> columns (matrix/float.hpp:504-512)
inline M4 fromRows(const float r[16])
{
	M4 m;
	for (int i = 0; i < 4; i++) for (int l = 0; l < 4; l++) m.c[i].v[l] = r[l * 4 + i];
	return m;
}

// inverse4x4 (simd/matrix/float.hpp:472-544): cofactor expansion over the transposed rows, in the reference's order of
// products, differences and fused steps; the determinant's reciprocal is one division (_mm_div_ss) scaled into every column.
M4 inverse(const M4& m)
{
	L4 t = shuf(m.c[0], m.c[1], 0, 1, 0, 1), r1 = shuf(m.c[2], m.c[3], 0, 1, 0, 1);
	const L4 r0 = shuf(t, r1, 0, 2, 0, 2);
	r1 = shuf(r1, t, 1, 3, 1, 3);
	t = shuf(m.c[0], m.c[1], 2, 3, 2, 3);
	L4 r3 = shuf(m.c[2], m.c[3], 2, 3, 2, 3);
	L4 r2 = shuf(t, r3, 0, 2, 0, 2);
	r3 = shuf(r3, t, 1, 3, 1, 3);
	auto swapPairs = [](const L4& x) { return shuf(x, x, 1, 0, 3, 2); };  // _MM_SHUFFLE(2, 3, 0, 1)
	auto swapHalves = [](const L4& x) { return shuf(x, x, 2, 3, 0, 1); }; // _MM_SHUFFLE(1, 0, 3, 2)

	t = swapPairs(mul(r2, r3));
	L4 m0 = mul(r1, t), m1 = mul(r0, t);
	t = swapHalves(t);
	m0 = sub(mul(r1, t), m0);
	m1 = swapHalves(sub(mul(r0, t), m1));

	t = swapPairs(mul(r1, r2));
	m0 = fma4(r3, t, m0);
	L4 m3 = mul(r0, t);
	t = swapHalves(t);
	m0 = sub(m0, mul(r3, t));
	m3 = swapHalves(sub(mul(r0, t), m3));

	t = swapPairs(mul(swapHalves(r1), r3));
	r2 = swapHalves(r2);
	m0 = fma4(r2, t, m0);
	L4 m2 = mul(r0, t);
	t = swapHalves(t);
	m0 = sub(m0, mul(r2, t));
	m2 = swapHalves(sub(mul(r0, t), m2));

	t = swapPairs(mul(r0, r1));
	m2 = fma4(r3, t, m2);
	m3 = sub(mul(r2, t), m3);
	t = swapHalves(t);
	m2 = sub(mul(r3, t), m2);
	m3 = sub(m3, mul(r2, t));

	t = swapPairs(mul(r0, r3));
	m1 = sub(m1, mul(r2, t));
	m2 = fma4(r1, t, m2);
	t = swapHalves(t);
	m1 = fma4(r2, t, m1);
	m2 = sub(m2, mul(r1, t));

	t = swapPairs(mul(r0, r2));
	m1 = fma4(r3, t, m1);
	m3 = sub(m3, mul(r1, t));
	t = swapHalves(t);
	m1 = sub(m1, mul(r3, t));
	m3 = fma4(r1, t, m3);

	L4 d = mul(r0, m0);
	d = add(swapPairs(d), d);
	const float det = rnd(swapHalves(d).v[0] + d.v[0]); // _mm_add_ss
	const L4 inv = splat(rnd(1.0f / det));
	M4 r;
	r.c[0] = mul(inv, m0); r.c[1] = mul(inv, m1); r.c[2] = mul(inv, m2); r.c[3] = mul(inv, m3);
	return r;
}

// lookAt(from, to, up = f32x4::top), matrix/transform.hpp:291-299
M4 lookAt(const L4& from, const L4& to)
{
	const L4 up = make(0.0f, 1.0f, 0.0f, 0.0f);
	const L4 f = normalize3(sub(to, from));
	const L4 s = normalize3(cross3(up, f));
	const L4 u = cross3(f, s);
	M4 rows;
	rows.c[0] = make(s.v[0], s.v[1], s.v[2], -dp3(s, from));
	rows.c[1] = make(u.v[0], u.v[1], u.v[2], -dp3(u, from));
	rows.c[2] = make(f.v[0], f.v[1], f.v[2], -dp3(f, from));
	rows.c[3] = make(0.0f, 0.0f, 0.0f, 1.0f);
	return transpose(rows);
}

M4 perspRevZ(float fov, float aspect, float nearPlane, float farPlane) // matrix/projection.hpp:56-66
{
	const float tanHalfFov = std::tan(rnd(fov * 0.5f));
	const float range = rnd(nearPlane - farPlane);
	const float rows[16] = { rnd(1.0f / rnd(aspect * tanHalfFov)), 0.0f, 0.0f, 0.0f,
		0.0f, rnd(-1.0f / tanHalfFov), 0.0f, 0.0f,
		0.0f, 0.0f, rnd(nearPlane / range), rnd(-rnd(nearPlane * farPlane) / range),
		0.0f, 0.0f, 1.0f, 0.0f };
	return fromRows(rows);
}
M4 perspInfRevZ(float fov, float aspect, float nearPlane) // matrix/projection.hpp:39-47
{
	const float tanHalfFov = std::tan(rnd(fov * 0.5f));
	const float rows[16] = { rnd(1.0f / rnd(aspect * tanHalfFov)), 0.0f, 0.0f, 0.0f,
		0.0f, rnd(-1.0f / tanHalfFov), 0.0f, 0.0f,
		0.0f, 0.0f, 0.0f, nearPlane,
		0.0f, 0.0f, 1.0f, 0.0f };
	return fromRows(rows);
}
M4 orthoRevZ(float w0, float w1, float h0, float h1, float d0, float d1) // matrix/projection.hpp:91-98
{
	const float w = rnd(w1 - w0), h = rnd(h1 - h0), d = rnd(d0 - d1);
	const float rows[16] = { rnd(2.0f / w), 0.0f, 0.0f, rnd(-rnd(w1 + w0) / w),
		0.0f, rnd(-2.0f / h), 0.0f, rnd(rnd(h1 + h0) / h),
		0.0f, 0.0f, rnd(1.0f / d), rnd(-d1 / d),
		0.0f, 0.0f, 0.0f, 1.0f };
	return fromRows(rows);
}

M4 load(const float* p) { M4 m; memcpy(&m, p, 64); return m; }
void store(float* p, const M4& m) { memcpy(p, &m, 64); }

// calcLightViewProj, source/system/render/csm.cpp:260-308
M4 lightViewProj(const M4& view, const L4& lightDir, L4& cameraOffset, float fov, float aspect, float nearPlane, float farPlane,
	float zCoeff, uint32_t shadowMapSize)
{
	const M4 invViewProj = inverse(mulMat(perspRevZ(fov, aspect, nearPlane, farPlane), view));
	L4 corners[8];
	int k = 0;
	for (int z = 0; z < 2; z++)
		for (int y = 0; y < 2; y++)
			for (int x = 0; x < 2; x++)
			{
				const L4 c = mulVec(invViewProj, make(rnd(x * 2.0f - 1.0f), rnd(y * 2.0f - 1.0f), (float)z, 1.0f));
				corners[k++] = dvd(c, splat(c.v[3]));
			}
	L4 centre = splat(0.0f);
	for (int i = 0; i < 8; i++)
		centre = add(centre, corners[i]);
	centre = mul(centre, splat(1.0f / 8.0f));
	const M4 lightView = lookAt(sub(centre, lightDir), centre);
	L4 mn = splat(3.402823466e+38f), mx = splat(-3.402823466e+38f); // f32x4::max / minusMax
	for (int i = 0; i < 8; i++)
	{
		const L4 trf = mulVec(lightView, corners[i]);
		mn = min4(mn, trf); mx = max4(mx, trf);
	}
	mn.v[2] = mn.v[2] < 0.0f ? rnd(mn.v[2] * zCoeff) : rnd(mn.v[2] / zCoeff);
	mx.v[2] = mx.v[2] < 0.0f ? rnd(mx.v[2] / zCoeff) : rnd(mx.v[2] * zCoeff);
	const float unitsPerTexel = rnd(rnd(mx.v[0] - mn.v[0]) / (float)shadowMapSize);
	L4 lightCameraPos = mulVec(lightView, centre);
	lightCameraPos.v[0] = rnd(std::floor(rnd(lightCameraPos.v[0] / unitsPerTexel)) * unitsPerTexel);
	lightCameraPos.v[2] = rnd(std::floor(rnd(lightCameraPos.v[2] / unitsPerTexel)) * unitsPerTexel);
	const L4 snapped = mulVec(inverse(lightView), lightCameraPos);
	const M4 stabilized = lookAt(sub(snapped, lightDir), snapped);
	cameraOffset = neg(add(mul(lightDir, splat(mn.v[2])), centre));
	return mulMat(orthoRevZ(mn.v[0], mx.v[0], mn.v[1], mx.v[1], mn.v[2], mx.v[2]), stabilized);
}

} // namespace

extern "C"
{

int gsp_light_view_proj(const float* view, const float* lightDir, float fieldOfView, float aspectRatio, float nearPlane, float farPlane,
	float zCoeff, uint32_t shadowMapSize, float* viewProj, float* cameraOffset)
{
	if (!view || !lightDir || !viewProj || !cameraOffset || shadowMapSize == 0)
		return GSP_ERR_INVALID;
	L4 offset;
	const M4 vp = lightViewProj(load(view), make(lightDir[0], lightDir[1], lightDir[2], 0.0f), offset, fieldOfView, aspectRatio,
		nearPlane, farPlane, zCoeff, shadowMapSize);
	store(viewProj, vp);
	memcpy(cameraOffset, offset.v, 16);
	return GSP_OK;
}

// Camera without a parent: view = rotate(normalize(q)) * translate(scale(s), -p) with its translation zeroed, projection =
// calcPerspProjInfRevZ, viewProj = projection * view (graphics.cpp:168-172,198-203,241; camera.hpp:111-121).
// rotate(normalize(q)), matrix/transform.hpp:128-140 (scalar float code on the normalised lanes)
static M4 rotationOf(const float* rotation)
{
	const L4 q = normalize4(make(rotation[0], rotation[1], rotation[2], rotation[3]));
	const float x = q.v[0], y = q.v[1], z = q.v[2], w = q.v[3];
	const float xx = rnd(x * x), yy = rnd(y * y), zz = rnd(z * z), xz = rnd(x * z), xy = rnd(x * y), yz = rnd(y * z);
	const float wx = rnd(w * x), wy = rnd(w * y), wz = rnd(w * z);
	M4 R;
	R.c[0] = make(rnd(1.0f - rnd(2.0f * rnd(yy + zz))), rnd(2.0f * rnd(xy + wz)), rnd(2.0f * rnd(xz - wy)), 0.0f);
	R.c[1] = make(rnd(2.0f * rnd(xy - wz)), rnd(1.0f - rnd(2.0f * rnd(xx + zz))), rnd(2.0f * rnd(yz + wx)), 0.0f);
	R.c[2] = make(rnd(2.0f * rnd(xz + wy)), rnd(2.0f * rnd(yz - wx)), rnd(1.0f - rnd(2.0f * rnd(xx + yy))), 0.0f);
	R.c[3] = make(0.0f, 0.0f, 0.0f, 1.0f);
	return R;
}
static M4 scaleOf(const float* scale) // scale(s), matrix/transform.hpp:80-84
{
	M4 S;
	S.c[0] = make(scale[0], 0.0f, 0.0f, 0.0f); S.c[1] = make(0.0f, scale[1], 0.0f, 0.0f);
	S.c[2] = make(0.0f, 0.0f, scale[2], 0.0f); S.c[3] = make(0.0f, 0.0f, 0.0f, 1.0f);
	return S;
}
// calcView (graphics.cpp:168-172): rotate(normalize(q)) * translate(scale(s), -p)
static M4 cameraView(const float* position, const float* rotation, const float* scale)
{
	const M4 R = rotationOf(rotation);
	M4 S = scaleOf(scale);
	const L4 t = neg(make(position[0], position[1], position[2], 0.0f));
	{
		// translate(m, t): c3 = (c3 + dot3x3(m, t)).xyz, lane W kept; dot3x3 = mul, fma, fma over the first three columns
		// (matrix/transform.hpp:59-62, simd/matrix/float.hpp:372-378)
		L4 r = mul(S.c[0], splat(t.v[0]));
		r = fma4(S.c[1], splat(t.v[1]), r);
		r = fma4(S.c[2], splat(t.v[2]), r);
		const L4 sum = add(S.c[3], r);
		S.c[3] = make(sum.v[0], sum.v[1], sum.v[2], S.c[3].v[3]);
	}
	return mulMat(R, S);
}
// math::calcModel, general branch (matrix/transform.hpp:255): translate(position) * rotate(normalize(q)) * scale(s). The
// `scale == f32x4::one` shortcut compares all four lanes and lane W of a component's scale holds childCapacity bits
// (transform.hpp:40,52,84), which never read as 1.0f: components always take this branch.
static M4 localModel(const float* position, const float* rotation, const float* scale)
{
	M4 T;
	T.c[0] = make(1.0f, 0.0f, 0.0f, 0.0f); T.c[1] = make(0.0f, 1.0f, 0.0f, 0.0f);
	T.c[2] = make(0.0f, 0.0f, 1.0f, 0.0f); T.c[3] = make(position[0], position[1], position[2], 1.0f);
	return mulMat(mulMat(T, rotationOf(rotation)), scaleOf(scale));
}

static int cameraViewProj(M4 V, float fieldOfView, float aspectRatio, float nearPlane, float* view, float* projection, float* viewProj)
{
	V.c[3] = make(0.0f, 0.0f, 0.0f, V.c[3].v[3]); // setTranslation(view, zero): the camera-relative view keeps lane W
	const M4 P = perspInfRevZ(fieldOfView, aspectRatio, nearPlane);
	store(view, V); store(projection, P); store(viewProj, mulMat(P, V));
	return GSP_OK;
}

int gsp_camera_view_proj(const float* position, const float* rotation, const float* scale, float fieldOfView, float aspectRatio,
	float nearPlane, float* view, float* projection, float* viewProj)
{
	if (!position || !rotation || !scale || !view || !projection || !viewProj)
		return GSP_ERR_INVALID;
	return cameraViewProj(cameraView(position, rotation, scale), fieldOfView, aspectRatio, nearPlane, view, projection, viewProj);
}

// calcRelativeView (graphics.cpp:173-189): view = calcView(camera); for every ancestor, nearest first:
// view = calcModel(ancestor) * view. parents[i] = { position xyz, rotation xyzw, scale xyz } (10 floats).
int gsp_camera_view_proj_chain(const float* position, const float* rotation, const float* scale, const float* parents,
	uint32_t parentCount, float fieldOfView, float aspectRatio, float nearPlane, float* view, float* projection, float* viewProj)
{
	if (!position || !rotation || !scale || !view || !projection || !viewProj || (parentCount && !parents))
		return GSP_ERR_INVALID;
	M4 V = cameraView(position, rotation, scale);
	for (uint32_t i = 0; i < parentCount; i++)
	{
		const float* a = parents + (size_t)i * 10;
		V = mulMat(localModel(a, a + 3, a + 7), V);
	}
	return cameraViewProj(V, fieldOfView, aspectRatio, nearPlane, view, projection, viewProj);
}

// The same for an ORTHOGRAPHIC CameraComponent (camera.hpp:119-120: calcOrthoProjRevZ(width, height, depth)); ancestors as in
// gsp_camera_view_proj_chain (parentCount may be 0).
int gsp_camera_view_proj_ortho(const float* position, const float* rotation, const float* scale, const float* parents,
	uint32_t parentCount, const float width[2], const float height[2], const float depth[2], float* view, float* projection,
	float* viewProj)
{
	if (!position || !rotation || !scale || !width || !height || !depth || !view || !projection || !viewProj || (parentCount && !parents))
		return GSP_ERR_INVALID;
	M4 V = cameraView(position, rotation, scale);
	for (uint32_t i = 0; i < parentCount; i++)
	{
		const float* a = parents + (size_t)i * 10;
		V = mulMat(localModel(a, a + 3, a + 7), V);
	}
	V.c[3] = make(0.0f, 0.0f, 0.0f, V.c[3].v[3]);
	const M4 P = orthoRevZ(width[0], width[1], height[0], height[1], depth[0], depth[1]);
	store(view, V); store(projection, P); store(viewProj, mulMat(P, V));
	return GSP_OK;
}

// The shadow passes of a frame the way CsmRenderSystem::prepareShadowRender produces them (csm.cpp:311-329): pass i covers
// [i == 0 ? cameraNear : distance * splits[i - 1],  i == count - 1 ? distance : distance * splits[i]].
int gsp_cascade_views(const float* view, const float* lightDir, float fieldOfView, float aspectRatio, float cameraNear,
	float shadowDistance, const float* splits, uint32_t cascadeCount, float zCoeff, uint32_t shadowMapSize, gsp_view* views,
	float* viewProjs)
{
	if (!view || !lightDir || !views || cascadeCount == 0 || cascadeCount > GSP_MAX_VIEWS || (cascadeCount > 1 && !splits))
		return GSP_ERR_INVALID;
	for (uint32_t i = 0; i < cascadeCount; i++)
	{
		float nearPlane = cameraNear, farPlane = shadowDistance;
		if (i > 0)
			nearPlane = rnd(shadowDistance * splits[i - 1]);
		if (i < cascadeCount - 1)
			farPlane = rnd(farPlane * splits[i]);
		float vp[16], offset[4];
		int rc = gsp_light_view_proj(view, lightDir, fieldOfView, aspectRatio, nearPlane, farPlane, zCoeff, shadowMapSize, vp, offset);
		if (rc) return rc;
		rc = gsp_view_from_viewproj(vp, offset, (int32_t)i, &views[i]);
		if (rc) return rc;
		if (viewProjs)
			memcpy(viewProjs + (size_t)i * 16, vp, 64);
	}
	return GSP_OK;
}

} // extern "C"
