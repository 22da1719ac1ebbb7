// kdpt_render.cuh — per-path device functions of the render loop: RNG, analytic geoms, scatterRay, shadeMaterial.
// Same operation order as the reference (and as oracle/kdpt_oracle.c); -fmad=false keeps + and * separate. The only
// places where results can differ from the reference's host build in the last ulp are the libm calls
// (sinf/cosf/acosf/powf/pow/cos/sin), which is why render-loop parity is stated with a tolerance (tests/).
#pragma once
#include "kdpt_core.cuh"
#include "../../include/kdpt.h"

namespace kdpt {

#define KDPT_PI 3.1415926535897932384626422832795028841971f
#define KDPT_TWO_PI 6.2831853071795864769252867665590057683943f
#define KDPT_SQRT_OF_ONE_THIRD 0.5773502691896257645091487805019574556476f

// utilhash, src/intersections.h:15-23
KDPT_HD unsigned utilhash(unsigned a)
{
    a = (a + 0x7ed55d16) + (a << 12);
    a = (a ^ 0xc761c23c) ^ (a >> 19);
    a = (a + 0x165667b1) + (a << 5);
    a = (a + 0xd3a2646c) ^ (a << 9);
    a = (a + 0xfd7046c5) + (a << 3);
    a = (a ^ 0xb55a4f09) ^ (a >> 16);
    return a;
}
// makeSeededRandomEngine, src/pathtrace.cu:62-66
KDPT_HD unsigned pathSeed(int iter, int index, int depth)
{
    return utilhash((1u << 31) | ((unsigned)depth << 22) | (unsigned)iter) ^ utilhash((unsigned)index);
}
// thrust::default_random_engine (minstd_rand: x <- 48271 x mod 2^31-1, seed 0 -> 1) and
// uniform_real_distribution<float>(0,1): float(x - 1) / (1.0f + float(2147483645))
struct Rng {
    unsigned x;
    // modulo 2^31-1 without a 64-bit division: 2^31 == 1 (mod m), so fold the high bits onto the low ones (exact)
    KDPT_HD void seed(unsigned s) { unsigned v = (s & 0x7fffffffu) + (s >> 31); if (v >= 2147483647u) v -= 2147483647u; x = v == 0 ? 1u : v; }
    KDPT_HD float u01()
    {
        unsigned long long p = (unsigned long long)x * 48271ull;
        unsigned v = (unsigned)(p & 0x7fffffffull) + (unsigned)(p >> 31);
        if (v >= 2147483647u) v -= 2147483647u;
        x = v;
        float r = (float)(x - 1u);
        r /= (1.0f + (float)(2147483646u - 1u));
        return (r * (1.0f - 0.0f)) + 0.0f;
    }
};

// glm mat4 * vec4 (type_mat4x4.inl:617-628): (m0*v0 + m1*v1) + (m2*v2 + m3*v3), column-major
KDPT_HD V3 mulMV(const float* m, V3 v, float w)
{
    return mk((m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * w),
              (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * w),
              (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * w));
}
KDPT_HD float gmin(float x, float y) { return (y < x) ? y : x; }
KDPT_HD float gmax(float x, float y) { return (x < y) ? y : x; }
KDPT_HD V3 pointOnRay(V3 o, V3 d, float t) { return o + (t - .0001f) * normalize(d); }

// boxIntersectionTest / sphereIntersectionTest, src/intersections.h:107-203. Returns t (-1 = miss); point, normal
// are written on a hit only.
KDPT_HD float geomTest(const kdpt_geom& g, V3 ro, V3 rd, V3& point, V3& normal)
{
    V3 qo = mulMV(g.inverseTransform, ro, 1.0f);
    V3 qd = normalize(mulMV(g.inverseTransform, rd, 0.0f));
    if (g.type == 1) {
        float tmin = -1e38f, tmax = 1e38f;
        V3 tminN = mk(0, 0, 0), tmaxN = mk(0, 0, 0);
#pragma unroll
        for (int xyz = 0; xyz < 3; ++xyz) {
            float qdxyz = comp(qd, xyz), qoxyz = comp(qo, xyz);
            float t1 = (-0.5f - qoxyz) / qdxyz;
            float t2 = (+0.5f - qoxyz) / qdxyz;
            float ta = gmin(t1, t2), tb = gmax(t1, t2);
            float s = t2 < t1 ? +1.0f : -1.0f;
            V3 n = mk(xyz == 0 ? s : 0.f, xyz == 1 ? s : 0.f, xyz == 2 ? s : 0.f);
            if (ta > 0 && ta > tmin) { tmin = ta; tminN = n; }
            if (tb < tmax) { tmax = tb; tmaxN = n; }
        }
        if (tmax >= tmin && tmax > 0) {
            if (tmin <= 0) { tmin = tmax; tminN = tmaxN; }
            point = mulMV(g.transform, pointOnRay(qo, qd, tmin), 1.0f);
            normal = normalize(mulMV(g.transform, tminN, 0.0f));
            return length(ro - point);
        }
        return -1.0f;
    } else if (g.type == 0) {
        float vDotDirection = dot(qo, qd);
        float radicand = vDotDirection * vDotDirection - (dot(qo, qo) - 0.25f); // powf(.5f, 2)
        if (radicand < 0) return -1.0f;
        float squareRoot = sqrtf(radicand);
        float firstTerm = -vDotDirection;
        float t1 = firstTerm + squareRoot, t2 = firstTerm - squareRoot;
        float t; bool outside;
        if (t1 < 0 && t2 < 0) return -1.0f;
        else if (t1 > 0 && t2 > 0) { t = fminf(t1, t2); outside = true; }
        else { t = fmaxf(t1, t2); outside = false; }
        V3 os = pointOnRay(qo, qd, t);
        point = mulMV(g.transform, os, 1.f);
        V3 n = normalize(mulMV(g.invTranspose, os, 0.f));
        normal = outside ? n : -n;
        return length(ro - point);
    }
    return -1.0f;
}

#if defined(__CUDACC__)
// The scatter helpers below are called from up to seven places of scatterRay. Inlined everywhere (KDPT_OUTLINE 0) they
// are ~40 % of the bounce kernel's 150-300 KB of SASS but ~3 % of its executed instructions, and that cold code is
// interleaved with the hot loops: ncu shows "no instruction" (instruction-cache miss) as the top or second stall reason
// of the bounce kernel. KDPT_OUTLINE 1 keeps ONE out-of-line copy of each (values in and out in registers, so the
// arithmetic — and every result bit — is the same).
#ifndef KDPT_OUTLINE
#define KDPT_OUTLINE 1
#endif
#if KDPT_OUTLINE >= 1
#define KDPT_SCATTER_FN __device__ __noinline__
#else
#define KDPT_SCATTER_FN __device__ __forceinline__
#endif
struct DirRng { float x, y, z; unsigned rng; };

// calculateRandomDirectionInHemisphere, src/interactions.h:9-41
KDPT_SCATTER_FN DirRng randomHemisphereImpl(float nx, float ny, float nz, unsigned rngState)
{
    Rng rng; rng.x = rngState;
    const V3 normal = mk(nx, ny, nz);
    float up = sqrtf(rng.u01());
    float over = sqrtf(1 - up * up);
    float around = rng.u01() * KDPT_TWO_PI;
    V3 dnn;
    if (fabsf(normal.x) < KDPT_SQRT_OF_ONE_THIRD) dnn = mk(1, 0, 0);
    else if (fabsf(normal.y) < KDPT_SQRT_OF_ONE_THIRD) dnn = mk(0, 1, 0);
    else dnn = mk(0, 0, 1);
    V3 p1 = normalize(cross(normal, dnn));
    V3 p2 = normalize(cross(normal, p1));
    const V3 d = (up * normal + (cosf(around) * over) * p1) + (sinf(around) * over) * p2;
    DirRng o; o.x = d.x; o.y = d.y; o.z = d.z; o.rng = rng.x;
    return o;
}
__device__ __forceinline__ V3 randomHemisphere(V3 normal, Rng& rng)
{
    const DirRng o = randomHemisphereImpl(normal.x, normal.y, normal.z, rng.x);
    rng.x = o.rng;
    return mk(o.x, o.y, o.z);
}
// rotateVector, src/interactions.h:44-65
__device__ __forceinline__ V3 rotateVector(V3 n1, V3 axis, float angle)
{
    axis = normalize(axis);
    float u = axis.x, v = axis.y, w = axis.z, x = n1.x, y = n1.y, z = n1.z;
    float ca = cosf(angle), sa = sinf(angle);
    return mk((-u * (-u * x - v * y - w * z)) * (1 - ca) + x * ca + (-w * y + v * z) * sa,
              (-v * (-u * x - v * y - w * z)) * (1 - ca) + y * ca + (w * x - u * z) * sa,
              (-w * (-u * x - v * y - w * z)) * (1 - ca) + z * ca + (-v * x + u * y) * sa);
}
// randSphericalVec, src/interactions.h:67-83 (theta, phi are doubles there)
__device__ __forceinline__ V3 randSpherical(float angle, Rng& rng)
{
    double theta = 2 * KDPT_PI * rng.u01();
    double phi = acosf((angle * KDPT_PI * rng.u01() - 1.0f));
    return normalize(mk((float)(cos(theta) * sin(phi)), (float)(sin(theta) * sin(phi)), (float)cos(phi)));
}
// getFresnelVal (Schlick), src/interactions.h:127-133
KDPT_SCATTER_FN float fresnelImpl(float ix, float iy, float iz, float nx, float ny, float nz, float ior)
{
    const V3 I = mk(ix, iy, iz), N = mk(nx, ny, nz);
    float R0 = powf((1.0f - ior) / (1.0f + ior), 2.0f);
    return (float)(R0 + (1.0f - R0) * pow((double)(1.0f - dot(N, -I)), 5.0));
}
__device__ __forceinline__ float fresnel(V3 I, V3 N, float ior) { return fresnelImpl(I.x, I.y, I.z, N.x, N.y, N.z, ior); }
__device__ __forceinline__ V3 reflectG(V3 I, V3 N) { return I - (N * dot(N, I)) * 2.0f; }
__device__ __forceinline__ V3 refractG(V3 I, V3 N, float eta)
{
    float dv = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - dv * dv);
    return (eta * I - (eta * dv + sqrtf(k)) * N) * (float)(k >= 0.0f);
}
KDPT_SCATTER_FN DirRng softJitterImpl(float dx, float dy, float dz, unsigned rngState, float a)
{
    Rng rng; rng.x = rngState;
    const V3 dir = mk(dx, dy, dz);
    V3 v = randSpherical(a, rng);
    float angle = acosf(dot(mk(0.0f, 0.0f, -1.0f), dir));
    V3 axis = normalize(cross(mk(0.0f, 0.0f, -1.0f), dir));
    const V3 d = rotateVector(v, axis, angle);
    DirRng o; o.x = d.x; o.y = d.y; o.z = d.z; o.rng = rng.x;
    return o;
}
__device__ __forceinline__ V3 softJitter(V3 dir, Rng& rng, float a)
{
    const DirRng o = softJitterImpl(dir.x, dir.y, dir.z, rng.x, a);
    rng.x = o.rng;
    return mk(o.x, o.y, o.z);
}

struct PathRay { V3 o, d; bool inside; float sdepth; };

// scatterRay, src/interactions.h:195-358
__device__ __forceinline__ void scatterRay(PathRay& r, V3 intersect, V3 normal, const kdpt_material& m, unsigned seed, float softness)
{
    Rng rng; rng.seed(seed);
    if (m.transmittance[0] > 0.0f || m.transmittance[1] > 0.0f || m.transmittance[2] > 0.0f) {
        float randval = rng.u01();
        if (randval < 0.5f && !r.inside) {
            r.d = softJitter(r.d, rng, 0.0001f);
            r.o = r.o + r.d * 0.0001f;
            r.sdepth = length(intersect - r.o);
            r.inside = true;
        } else {
            r.d = randomHemisphere(normal, rng);
            r.o = intersect + normal * 0.00001f;
            r.sdepth = 0.0f;
        }
    } else if (m.hasRefractive != 0.0f) {
        float randval = rng.u01();
        r.d = normalize(r.d);
        normal = normalize(normal);
        float fresn = fresnel(r.d, normal, m.indexOfRefraction);
        if (randval < 1.0f - fresn) {
            float ior = m.indexOfRefraction;
            if (!r.inside) ior = 1.0f / m.indexOfRefraction;
            float angle = (float)(1.0f - pow((double)ior, 2.0) * (1.0f - pow((double)dot(normal, r.d), 2.0)));
            if (angle < 0.0f) {
                float val = rng.u01();
                if (val < m.hasReflective) {
                    r.d = reflectG(r.d, normal);
                    if (softness > 0.0f) r.d = softJitter(r.d, rng, 0.02f);
                    r.o = intersect + normal * 0.00001f;
                } else {
                    r.d = randomHemisphere(normal, rng);
                    r.o = intersect + normal * 0.00001f;
                }
            } else {
                float val = rng.u01();
                if (val < m.hasRefractive) {
                    r.d = refractG(r.d, normal, ior);
                    if (softness > 0.0f) r.d = softJitter(r.d, rng, 0.02f);
                    r.o = intersect - normal * 0.001f;
                    r.inside = !r.inside;
                } else {
                    r.d = randomHemisphere(normal, rng);
                    r.o = intersect + normal * 0.00001f;
                }
            }
        } else {
            r.d = reflectG(r.d, normal);
            r.o = intersect + normal * 0.00001f;
            r.inside = false;
        }
    } else if (m.hasReflective != 0.0f) {
        float randval = rng.u01();
        if (randval < m.hasReflective) {
            r.d = reflectG(r.d, normal);
            if (softness > 0.0f) r.d = softJitter(r.d, rng, 0.02f);
            r.o = intersect + normal * 0.0001f;
            r.inside = false;
        } else {
            r.d = randomHemisphere(normal, rng);
            r.o = intersect + normal * 0.00001f;
        }
    } else {
        r.d = randomHemisphere(normal, rng);
        r.o = intersect + normal * 0.00001f;
        r.inside = false;
    }
}

// shadeMaterial for a path that hit something (t > 0), src/pathtrace.cu:2326-2361. Returns the new remainingBounces.
__device__ __forceinline__ int shadeHit(const kdpt_material& m, bool enableSss, float sdepth, V3& color, int remaining)
{
    V3 mc = mk(m.color[0], m.color[1], m.color[2]);
    if (m.emittance > 0.0f) {
        color = color * (mc * m.emittance);
        return 0;
    }
    V3 sc = mk(m.specular_color[0], m.specular_color[1], m.specular_color[2]);
    V3 tr = mk(m.transmittance[0], m.transmittance[1], m.transmittance[2]);
    if (enableSss && (tr.x > 0.0f || tr.y > 0.0f || tr.z > 0.0f)) {
        float sss = (double)(1.0f * sdepth) > 1.0 ? 1.0f : sdepth;
        sss = (double)(1.0f - sss) < 0.0 ? 0.0f : sss;
        sss = (float)pow((double)sss, 2.0);
        color = color * ((mc * 1.0f + m.hasRefractive * sc) + sss * tr);
    } else if (m.hasRefractive > 0.0f) color = color * (mc * 1.0f + m.hasRefractive * sc);
    else if (m.hasReflective > 0.0f) color = color * (mc * 1.0f + m.hasReflective * sc);
    else color = color * (mc * 1.0f);
    return remaining - 1;
}
#endif // __CUDACC__

} // namespace kdpt
