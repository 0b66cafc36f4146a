// bwd_common.cuh — epilogue math shared by the two backward sweep kernels: one S tile slice (this thread's row, 64
// columns) -> dS values, staged as scaled fp16.
#pragma once
#include "clip_kernels.cuh"
#include "sm100.cuh"
#include <cuda_fp16.h>

namespace flyp {

// Power-of-two staging scale: |dS| * G < 2^14 with |dS| <= max|g| (bit pattern of the max in *gmax_bits).
__device__ __forceinline__ void staging_scale(const uint32_t* gmax_bits, float& G, float& invG) {
    G = 1.f; invG = 1.f;
    const uint32_t gb = *gmax_bits;
    if ((gb & 0x7fffffffu) != 0u) {
        int ge = 13 - ((int)((gb >> 23) & 0xffu) - 127);
        ge = ge < -100 ? -100 : (ge > 100 ? 100 : ge);
        G = __uint_as_float((uint32_t)(ge + 127) << 23);
        invG = __uint_as_float((uint32_t)(127 - ge) << 23);
    }
}

// Per-row constants of the epilogue thread.
struct RowCtx {
    float w;      // G * wr[m]                       (robust form)
    float l;      // lr[m], log2 units               (robust form)
    float a;      // G * fa[m] = G wr 2^(c0 - lr)    (fast form)
    float d;      // G * dr[m]: exact dS at the positive
    int lab;      // positive column or -1
    int m;        // global row
    int cls;      // class id of the row (label-aware variants)
    float k;      // G * mk_r[m]                       (label-aware variants)
};

template <bool ROW_TERM>
__device__ __forceinline__ RowCtx load_row_ctx(const BwdParams& p, int m, bool fast, float G) {
    RowCtx c;
    c.w = 0.f; c.l = 0.f; c.a = 0.f; c.d = 0.f; c.lab = -1; c.m = m; c.cls = -2; c.k = 0.f;
    if (m < p.n_m) {
        if (p.mask_mode != 0) { c.cls = p.cls_m[m]; c.k = p.mk_r != nullptr ? p.mk_r[m] * G : 0.f; }
        if (ROW_TERM) {
            if (fast) c.a = p.fa[m] * G;
            else { c.w = p.wr[m] * G; c.l = p.lr[m]; }
        }
        if (p.labr != nullptr) { c.lab = p.labr[m]; c.d = p.dr[m] * G; }
    }
    return c;
}

// r0/r1: raw dot products of this thread's row with columns [n0, n0 + 64).  v: dS * G.  dsum += sum v * dot.
//   robust: dS = wr 2^(x - lr) + wc 2^(x - lc)                      (two exponentials per element)
//   fast:   dS = 2^(x - c0) * (fa[m] + fb[n]),  fa = wr 2^(c0 - lr), fb = wc 2^(c0 - lc)   (one exponential);
//           valid when every |lse - c0| <= 100 (checked by k_bwd_fast_vectors), then neither factor over/underflows
//           in a way that matters: a flushed 2^(x - c0) corresponds to softmax weights below 2^-26.
template <bool ROW_TERM, bool COL_TERM, bool MASK = false>
__device__ __forceinline__ void ds_tile(const uint32_t (&r0)[32], const uint32_t (&r1)[32], const BwdParams& p,
                                        const RowCtx& rc, int n0, float c1, bool fast, float c0, float G,
                                        float (&v)[64], bool want_ds, float& dsum) {
    if (fast) {
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
            float fb[4] = {0.f, 0.f, 0.f, 0.f};
            if (COL_TERM) {
                const float4 f4 = __ldg(reinterpret_cast<const float4*>(p.fb + n0) + k4);
                fb[0] = fmaf(f4.x, G, rc.a); fb[1] = fmaf(f4.y, G, rc.a);
                fb[2] = fmaf(f4.z, G, rc.a); fb[3] = fmaf(f4.w, G, rc.a);
            } else {
                fb[0] = fb[1] = fb[2] = fb[3] = rc.a;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k4 * 4 + j;
                const float e = sm100::ex2f(fmaf(__uint_as_float(k < 32 ? r0[k & 31] : r1[k & 31]), c1, -c0));
                v[k] = e * fb[j];
            }
        }
    } else {
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
            float lcs[4] = {0.f, 0.f, 0.f, 0.f}, wcs[4] = {0.f, 0.f, 0.f, 0.f};
            if (COL_TERM) {
                const float4 lc4 = __ldg(reinterpret_cast<const float4*>(p.lc + n0) + k4);
                const float4 wc4 = __ldg(reinterpret_cast<const float4*>(p.wc + n0) + k4);
                lcs[0] = lc4.x; lcs[1] = lc4.y; lcs[2] = lc4.z; lcs[3] = lc4.w;
                wcs[0] = wc4.x * G; wcs[1] = wc4.y * G; wcs[2] = wc4.z * G; wcs[3] = wc4.w * G;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k4 * 4 + j;
                const float x = __uint_as_float(k < 32 ? r0[k & 31] : r1[k & 31]) * c1;
                float acc = 0.f;
                if (ROW_TERM) acc = rc.w * sm100::ex2f(x - rc.l);
                if (COL_TERM) acc = fmaf(wcs[j], sm100::ex2f(x - lcs[j]), acc);
                v[k] = acc;
            }
        }
    }
    if (MASK) {
        // label-aware variants: the entries whose column has the row's class (the positive itself is substituted below)
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
            const int4 c4 = __ldg(reinterpret_cast<const int4*>(p.cls_n + n0) + k4);
            const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
            float kcs[4] = {0.f, 0.f, 0.f, 0.f}, lcs[4] = {0.f, 0.f, 0.f, 0.f};
            if (p.mask_mode >= 2) {
                const float4 k4v = __ldg(reinterpret_cast<const float4*>(p.mk_c + n0) + k4);
                kcs[0] = k4v.x * G; kcs[1] = k4v.y * G; kcs[2] = k4v.z * G; kcs[3] = k4v.w * G;
            }
            if (p.mask_mode == 3) {
                const float4 l4 = __ldg(reinterpret_cast<const float4*>(p.lc + n0) + k4);
                lcs[0] = l4.x; lcs[1] = l4.y; lcs[2] = l4.z; lcs[3] = l4.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k4 * 4 + j;
                const bool same = cc[j] == rc.cls;
                float nv;
                if (p.mask_mode == 1) nv = 0.f;
                else if (p.mask_mode == 2) nv = v[k] - (rc.k + kcs[j]);
                else {
                    const float x = __uint_as_float(k < 32 ? r0[k & 31] : r1[k & 31]) * c1;
                    nv = v[k] - (__fdividef(rc.k, 1.f - sm100::ex2f(x - rc.l)) + __fdividef(kcs[j], 1.f - sm100::ex2f(x - lcs[j])));
                }
                v[k] = same ? nv : v[k];
            }
        }
    }
    if (p.labr != nullptr) {
        const int rel = rc.lab - n0;
        if (__any_sync(0xffffffffu, rel >= 0 && rel < 64)) {
#pragma unroll
            for (int k = 0; k < 64; ++k) v[k] = (k == rel) ? rc.d : v[k];
        }
    }
    if (p.labc != nullptr) {
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
            const int4 lb4 = __ldg(reinterpret_cast<const int4*>(p.labc + n0) + k4);
            const float4 dc4 = __ldg(reinterpret_cast<const float4*>(p.dc + n0) + k4);
            v[k4 * 4 + 0] = (lb4.x == rc.m) ? dc4.x * G : v[k4 * 4 + 0];
            v[k4 * 4 + 1] = (lb4.y == rc.m) ? dc4.y * G : v[k4 * 4 + 1];
            v[k4 * 4 + 2] = (lb4.z == rc.m) ? dc4.z * G : v[k4 * 4 + 2];
            v[k4 * 4 + 3] = (lb4.w == rc.m) ? dc4.w * G : v[k4 * 4 + 3];
        }
    }
    // d(loss)/d(scale) = sum dS * <a, b>: accumulated here from the raw dot products (zero for padded rows / columns)
    if (want_ds) {
#pragma unroll
        for (int k = 0; k < 64; ++k)
            dsum = fmaf(v[k], __uint_as_float(k < 32 ? r0[k & 31] : r1[k & 31]), dsum);
    }
}

__device__ __forceinline__ void pack_ds(const float (&v)[64], uint32_t (&pk)[32]) {
#pragma unroll
    for (int k = 0; k < 32; ++k) pk[k] = sm100::pack_f16x2(v[2 * k], v[2 * k + 1]);
}
// v -= float(fp16(v)) given the packed fp16 pairs of v
__device__ __forceinline__ void residual_ds(float (&v)[64], const uint32_t (&pk)[32]) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const __half2 hh = *reinterpret_cast<const __half2*>(&pk[k]);
        const float2 f = __half22float2(hh);
        v[2 * k] -= f.x;
        v[2 * k + 1] -= f.y;
    }
}
// store this thread's 64 staged values as one 128-byte row segment of a K-major SWIZZLE_128B chunk
__device__ __forceinline__ void store_ds_row(uint8_t* chunk_row, int rloc, const uint32_t (&pk)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint4 val = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        *reinterpret_cast<uint4*>(chunk_row + ((j ^ (rloc & 7)) << 4)) = val;
    }
}

}  // namespace flyp
