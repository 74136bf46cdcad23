// dynavsr_b200/csrc/tc_common.cuh -- inline-PTX wrappers for the Blackwell tensor-core path:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05.mma / commit / ld, TMEM allocation, shared-memory matrix
// descriptors.  Bit layouts follow cute/arch/mma_sm100_desc.hpp (CUTLASS headers vendored in this image);
// the two non-obvious hardware facts used here were measured with tools/umma_probe.cu on a B200:
//   (1) SWIZZLE_128B descriptors may start at ANY 128-byte row of a TMA-written tile with base_offset = 0 and
//       a stride-byte-offset that is not a multiple of 1024 (the MMA swizzles on absolute smem address bits,
//       exactly like TMA) -> tap-shifted convolution windows can be served from one halo tile;
//   (2) MN-major TF32 operands need the 32-byte-atom swizzle: TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, descriptor
//       layout type 1 (SWIZZLE_128B_BASE32B), LBO = byte stride between 32-element MN blocks, SBO = 512 (next
//       4 K-rows); an M = 64 accumulator keeps row m in TMEM lane (m / 16) * 32 + m % 16.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace dvsr {

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t a = smem_u32(bar);
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    }
}
// Same wait with a sleep between polls, for warps that wait long and are not on the critical path (epilogue, producer, MMA
// issuer of a gather-bound kernel): a tight try_wait / branch loop competes for issue slots with the working warps of the
// same scheduler (ncu: 40 % of the instructions of the DCN kernel were such polls).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns = 128) {
    uint32_t done = 0;
    const uint32_t a = smem_u32(bar);
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 (1024 B between
//   8-row groups) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6),
// a/b_format TF32 (2) at [7,10)/[10,13), a/b major K (0) at 15/16, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 with BF16 operands, fp32 accumulation (a/b_format = 1)
__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// x = hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi); returns the two bf16 bit patterns (x - hi - lo ~ 2^-17 |x|)
__device__ __forceinline__ void split_bf16(float x, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = (uint32_t)__bfloat16_as_ushort(h);
    lo = (uint32_t)__bfloat16_as_ushort(l);
}
// the same split for two values at once with the packed converters (F2FP.BF16.PACK_AB): hi = {bf16(x1), bf16(x0)},
// lo = {bf16(x1 - hi1), bf16(x0 - hi0)}; x0 lands in the low half-word (element order of a K-major row)
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float r0 = x0 - __uint_as_float(hi << 16);
    const float r1 = x1 - __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// the same split on a float2 with packed fp32x2 math (sm_100 FFMA2): the residual of both lanes in one instruction
__device__ __forceinline__ void split_bf16x2_packed(float2 x, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __float22bfloat162_rn(x);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float2 hf = make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
    const float2 r = __ffma2_rn(hf, make_float2(-1.f, -1.f), x);
    const __nv_bfloat162 l = __float22bfloat162_rn(r);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// explicit shared-window accesses (a pointer obtained by integer arithmetic on the dynamic shared-memory base is generic to
// the compiler: it emits LD.E / ST.E with a generic-address check instead of LDS / STS)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}


// generic descriptor: layout type 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (MN-major tf32)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t ltype) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)ltype << 61;
    return d;
}
// kind::tf32 instruction descriptor with explicit major-ness (0 = K-major, 1 = MN-major)
__device__ __forceinline__ uint32_t make_idesc_tf32_major(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 256-bit global store (STG.E.256, sm_100+): p must be 32-byte aligned
__device__ __forceinline__ void st_global_v8(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h) : "memory");
}

// Epilogue math for one 32-column accumulator chunk held in registers (channels c0 .. c0+31 of one pixel):
//   v = act(v + addend + bias) + res.  `bias32` points at 32 consecutive bias values (shared or global memory,
//   16-byte aligned) or is null; addend / res point at this pixel's channel c0 (16-byte aligned) or are null.
//   Channels >= Co are left untouched (they are never stored).
__device__ __forceinline__ void epilogue_chunk(float (&v)[32], int c0, int Co, const float* bias32, const float* addend,
                                               const float* res, int act, float slope, int sig_split) {
    const int nvalid = min(32, Co - c0);          // multiple of 4
    if (addend) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            if (j < nvalid) { const float4 t = ldg4(addend + j); v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w; }
    }
    if (bias32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            if (j < nvalid) {
                const float4 t = *reinterpret_cast<const float4*>(bias32 + j);
                v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
    }
    if (act == DVSR_ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * slope;
    } else if (act == DVSR_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == DVSR_ACT_SIGMOID_SPLIT && c0 + 32 > sig_split) {
        // only the chunks that reach into the mask channels pay for the sigmoid; ex2 + fast reciprocal (2 MUFU ops) instead
        // of an IEEE division per element -- the epilogue of the 64 -> 216 offset/mask conv was its bottleneck
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c0 + j >= sig_split && j < nvalid) v[j] = __fdividef(1.f, 1.f + __expf(-v[j]));
    }
    if (res) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            if (j < nvalid) { const float4 t = ldg4(res + j); v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w; }
    }
}

// ------------------------------------------------------------------------------------------------ coalesced epilogue stores
// tcgen05.ld.32x32b leaves one accumulator ROW (pixel) per thread: storing it directly makes every warp-level store touch 32
// different 128-byte lines (one 32-byte sector each).  The round-2 ncu capture of conv_tc2 showed what that costs: 59 L1
// data-stage wavefronts per STG.256 request, 24 % of the (saturated) L1 / shared-memory data pipe of the kernel.  A 4 x 4
// transpose of 8-float pieces inside each lane quad (2 butterfly steps, 32 SHFL + 96 SEL per thread) turns "32 channels of my
// pixel" into "my 8 channels of the quad's 4 consecutive pixels", so that in each store instruction a quad writes one full
// 128-byte line (8 lines per warp instruction instead of 32 sectors).
//   in : v[8 * c + e] = channel (c0 + 8 * c + e) of pixel row `lane`            (c = 0..3, e = 0..7)
//   out: v[8 * j + e] = channel (c0 + 8 * (lane & 3) + e) of pixel row (lane & ~3) + j
__device__ __forceinline__ void quad_transpose32(float (&v)[32], int lane) {
    const bool b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float x = b1 ? v[8 * s + e] : v[8 * (s + 2) + e];
            x = __shfl_xor_sync(0xffffffffu, x, 2);
            if (b1) v[8 * s + e] = x; else v[8 * (s + 2) + e] = x;
        }
#pragma unroll
    for (int s = 0; s < 4; s += 2)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float x = b0 ? v[8 * s + e] : v[8 * (s + 1) + e];
            x = __shfl_xor_sync(0xffffffffu, x, 1);
            if (b0) v[8 * s + e] = x; else v[8 * (s + 1) + e] = x;
        }
}
// Epilogue in the transposed domain: v[8 * j + e] = channel (cch + e) of pixel (pixb + j), j < nok (the quad's 4 consecutive
// pixels, the first nok of them inside the image):  y = act(v + addend + bias) + res, stored as 32 contiguous bytes per pixel
// (one full line per quad).  bias8 -> 8 bias values of channels cch.. (shared or global, 16-byte aligned) or null; addend / res /
// y are tensor bases (16-byte aligned rows; vec8: y rows 32-byte aligned).  Channels >= Co are neither read nor written
// (Co is a multiple of 4).
__device__ __forceinline__ void epilogue_store_t(float (&v)[32], const float* bias8, int cch, int Co, long long pixb, int nok,
                                                 const float* addend, int addend_pix_stride, const float* res, int res_pix_stride,
                                                 float* y, int y_pix_stride, bool vec8, int act, float slope, int sig_split,
                                                 bool dry = false) {
    // dry: run every instruction of the body with the loads and stores predicated off (instruction-cache warm-up pass of a
    // persistent kernel's epilogue warps while they wait for their first accumulator)
    const int nch = Co - cch;
    if (nch <= 0) return;
    const bool full = nch >= 8;
    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
    if (bias8) { b0 = *reinterpret_cast<const float4*>(bias8); b1 = *reinterpret_cast<const float4*>(bias8 + 4); }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < nok || dry) {
            float* w = &v[8 * j];
            const long long px = pixb + j;
            if (addend && !dry) {
                const float* a = addend + px * addend_pix_stride + cch;
                const float4 t0 = ldg4(a);
                w[0] += t0.x; w[1] += t0.y; w[2] += t0.z; w[3] += t0.w;
                if (full) { const float4 t1 = ldg4(a + 4); w[4] += t1.x; w[5] += t1.y; w[6] += t1.z; w[7] += t1.w; }
            }
            w[0] += b0.x; w[1] += b0.y; w[2] += b0.z; w[3] += b0.w; w[4] += b1.x; w[5] += b1.y; w[6] += b1.z; w[7] += b1.w;
            if (act == DVSR_ACT_LRELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = w[e] > 0.f ? w[e] : w[e] * slope;
            } else if (act == DVSR_ACT_RELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = fmaxf(w[e], 0.f);
            } else if (act == DVSR_ACT_SIGMOID_SPLIT && cch + 8 > sig_split) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (cch + e >= sig_split) w[e] = __fdividef(1.f, 1.f + __expf(-w[e]));
            }
            if (res && !dry) {
                const float* a = res + px * res_pix_stride + cch;
                const float4 t0 = ldg4(a);
                w[0] += t0.x; w[1] += t0.y; w[2] += t0.z; w[3] += t0.w;
                if (full) { const float4 t1 = ldg4(a + 4); w[4] += t1.x; w[5] += t1.y; w[6] += t1.z; w[7] += t1.w; }
            }
            float* yo = y + px * y_pix_stride + cch;
            if (dry) {
                // nothing leaves the thread
            } else if (vec8 && full) {
                st_global_v8(yo, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
            } else {
                *reinterpret_cast<float4*>(yo) = make_float4(w[0], w[1], w[2], w[3]);
                if (full) *reinterpret_cast<float4*>(yo + 4) = make_float4(w[4], w[5], w[6], w[7]);
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

}  // namespace dvsr
