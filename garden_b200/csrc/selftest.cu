// Device self-test of the exact-arithmetic shortcuts (diagnostics entry point gsp_selftest_math, used by tests/).
//
// The hot kernel replaces IEEE divisions / square roots / 4-lane products by cheaper instruction sequences that are
// bit-identical inside explicit guards (sceneprep_math.cuh: localModel43Fast, matMul43P). Whether they are depends on what the
// hardware's approximation units return, so the claim is checked ON the device: random and adversarial inputs run through
// the shortcut and through the long form (the reference's operation order, 4 lanes, IEEE intrinsics), results compared bit
// for bit.
#include "sceneprep_internal.h"
#include "sceneprep_math.cuh"

namespace gsp
{

__device__ __forceinline__ uint64_t mix64(uint64_t& s)
{
	s += 0x9E3779B97F4A7C15ull;
	uint64_t z = s;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t& s) { return (float)(mix64(s) >> 40) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float weird(uint64_t& s, uint32_t mode)
{
	const uint64_t r = mix64(s);
	switch (mode & 7u)
	{
	case 0: return 0.0f;
	case 1: return -0.0f;
	case 2: return __uint_as_float((uint32_t)r);                                  // any bit pattern (NaN, Inf, subnormal, ...)
	case 3: return __uint_as_float(((uint32_t)r & 0x807fffffu) | (1u << 23));     // smallest normals
	case 4: return __uint_as_float((uint32_t)r & 0x807fffffu);                    // subnormals
	case 5: return ldexpf(u01(s) * 2.0f - 1.0f, (int)(r % 141) - 70);             // 2^-70 .. 2^70
	case 6: return (r & 1) ? 1.0f : -1.0f;
	default: return u01(s) * 2.0f - 1.0f;
	}
}

__global__ void kSelftestMath(uint64_t seed, uint32_t iterations, unsigned long long* out /* tested, fast, mismatch, mm tested, mm mismatch */)
{
	uint64_t s = seed ^ ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0xD1342543DE82EF95ull);
	unsigned long long tested = 0, fastTaken = 0, mismatches = 0, mmTested = 0, mmMismatches = 0;
	Mat43 prev;
	for (int i = 0; i < 4; i++) for (int l = 0; l < 3; l++) prev.c[i][l] = (i == l) ? 1.0f : 0.0f;
	for (uint32_t it = 0; it < iterations; it++)
	{
		float q[4], sc[3], p[3];
		const uint32_t kind = (uint32_t)(mix64(s) & 15u);
		// unit-ish quaternion, optionally scaled by a power of two, optionally with special components
		float len2 = 0.0f;
		for (int k = 0; k < 4; k++) { q[k] = u01(s) * 2.0f - 1.0f; len2 += q[k] * q[k]; }
		const float inv = rsqrtf(fmaxf(len2, 1e-12f)) * (0.9f + 0.2f * u01(s));
		for (int k = 0; k < 4; k++) q[k] *= inv;
		for (int k = 0; k < 3; k++) { sc[k] = 0.25f + 1.75f * u01(s); p[k] = (u01(s) - 0.5f) * 4000.0f; }
		if (kind >= 8)
		{
			if (kind == 8) { const float f = ldexpf(1.0f, (int)(mix64(s) % 101) - 50); for (int k = 0; k < 4; k++) q[k] *= f; }
			if (kind == 9 || kind == 10) { const uint64_t z = mix64(s); for (int k = 0; k < 4; k++) if ((z >> k) & 1) q[k] = (z >> (8 + k)) & 1 ? -0.0f : 0.0f; }
			if (kind == 11) q[mix64(s) & 3] = weird(s, (uint32_t)mix64(s));
			if (kind == 12) for (int k = 0; k < 4; k++) q[k] = weird(s, (uint32_t)mix64(s));
			if (kind == 13) sc[mix64(s) % 3] = weird(s, (uint32_t)mix64(s));
			if (kind == 14) { sc[0] = 1.0f; sc[1] = -1.0f; sc[2] = 1.0f; p[mix64(s) % 3] = weird(s, (uint32_t)mix64(s)); }
			if (kind == 15) { const float f = ldexpf(1.0f, (int)(mix64(s) % 241) - 120); for (int k = 0; k < 3; k++) sc[k] *= f; }
		}
		Mat43 fastL;
		const bool ok = localModel43Fast<true>(p[0], p[1], p[2], q[0], q[1], q[2], q[3], sc[0], sc[1], sc[2], fastL);
		const Mat4 exact = localModel(p[0], p[1], p[2], q[0], q[1], q[2], q[3], sc[0], sc[1], sc[2]);
		tested++;
		if (ok)
		{
			fastTaken++;
			bool same = true;
			for (int i = 0; i < 4; i++)
				for (int l = 0; l < 3; l++)
					same = same && (__float_as_uint(fastL.c[i][l]) == __float_as_uint(exact.c[i][l]));
			if (!same) mismatches++;
			// packed product against the scalar 4-lane product, chained so that operands look like real world matrices
			bool finite = true, prevFinite = true;
			for (int i = 0; i < 4; i++) for (int l = 0; l < 3; l++)
			{
				finite = finite && fabsf(fastL.c[i][l]) < 1e15f;
				prevFinite = prevFinite && fabsf(prev.c[i][l]) < 1e15f;
			}
			if (finite && !prevFinite)
			{
				prev = fastL; prevFinite = true;
			}
			if (finite && prevFinite)
			{
				Mat4 a4, b4;
				for (int i = 0; i < 4; i++) { for (int l = 0; l < 3; l++) { a4.c[i][l] = fastL.c[i][l]; b4.c[i][l] = prev.c[i][l]; } a4.c[i][3] = b4.c[i][3] = (i == 3) ? 1.0f : 0.0f; }
				const Mat4 want = matMul(a4, b4);
				const float4 rows[3] = { make_float4(fastL.c[0][0], fastL.c[1][0], fastL.c[2][0], fastL.c[3][0]),
					make_float4(fastL.c[0][1], fastL.c[1][1], fastL.c[2][1], fastL.c[3][1]),
					make_float4(fastL.c[0][2], fastL.c[1][2], fastL.c[2][2], fastL.c[3][2]) };
				const Mat43 got = unpairMat(matMul43P(rows, pairMat(prev)));
				bool sameMM = true;
				for (int i = 0; i < 4; i++) for (int l = 0; l < 3; l++) sameMM = sameMM && (__float_as_uint(got.c[i][l]) == __float_as_uint(want.c[i][l]));
				mmTested++;
				if (!sameMM) mmMismatches++;
				if ((it & 7u) == 7u) prev = fastL; else prev = got; // chains of up to 8 products, then restart
			}
		}
	}
	atomicAdd(&out[0], tested); atomicAdd(&out[1], fastTaken); atomicAdd(&out[2], mismatches);
	atomicAdd(&out[3], mmTested); atomicAdd(&out[4], mmMismatches);
}

} // namespace gsp

extern "C" int gsp_selftest_math(int device, uint32_t blocks, uint32_t iterations, uint64_t seed, uint64_t results[5])
{
	if (!results || blocks == 0)
		return GSP_ERR_INVALID;
	if (cudaSetDevice(device) != cudaSuccess)
		return GSP_ERR_CUDA;
	unsigned long long* d = nullptr;
	if (cudaMalloc((void**)&d, 5 * sizeof(unsigned long long)) != cudaSuccess)
		return GSP_ERR_NOMEM;
	cudaMemset(d, 0, 5 * sizeof(unsigned long long));
	gsp::kSelftestMath<<<blocks, 256>>>(seed, iterations, d);
	cudaError_t err = cudaMemcpy(results, d, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
	cudaFree(d);
	return err == cudaSuccess ? GSP_OK : GSP_ERR_CUDA;
}
