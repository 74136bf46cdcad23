// tools/umma_probe.cu -- hardware-semantics probes for two tcgen05 descriptor questions that the guides
// do not answer (no public docs in this sandbox):
//   P1: can a K-major SWIZZLE_128B A-descriptor start at an arbitrary 128-byte row of a TMA-written halo
//       tile, with a stride-byte-offset that is not a multiple of 1024 (shifted conv windows served from
//       ONE halo tile instead of 9 re-loads)?  Tried with base_offset = 0 and base_offset = (addr>>7)&7.
//   P2: MN-major SWIZZLE_128B operands (weight-gradient GEMM straight from NHWC tiles): which of LBO / SBO
//       is the stride between 32-element MN blocks, and where do the rows of an M=64 / M=128 accumulator
//       land in TMEM?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma3(const CUtensorMap* m, uint64_t* bar, void* dst, int a, int b, int c) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(s32(dst)), "l"(m), "r"(s32(bar)), "r"(a), "r"(b), "r"(c) : "memory");
}
__device__ __forceinline__ void tma2(const CUtensorMap* m, uint64_t* bar, void* dst, int a, int b) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(dst)), "l"(m), "r"(s32(bar)), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory"); }
__device__ __forceinline__ void ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off, uint32_t ltype = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)ltype << 61;
    return d;
}

// ---------------------------------------------------------------------------------------------- P1
// halo tile [18][10][32] (TMA, SW128) ; window (dy, dx) of 16 x 8 pixels ; B = [64][32] K-major
__global__ void probe1(const __grid_constant__ CUtensorMap mh, const __grid_constant__ CUtensorMap mw, int dy, int dx, int mode, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = sm;                 // 18*10*128 = 23040 B
    uint8_t* sb = sm + 23552;         // 64*128 = 8192 B
    uint64_t* bar = (uint64_t*)(sb + 8192);
    uint32_t* slot = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    if (threadIdx.x == 0) {
        mbar_expect(&bar[0], 23040 + 8192);
        tma3(&mh, &bar[0], sa, 0, 0, 0);
        tma2(&mw, &bar[0], sb, 0, 0);
        mbar_wait(&bar[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = s32(sa) + (dy * 10 + dx) * 128;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        for (int k = 0; k < 4; ++k) {
            const uint32_t bo = mode == 0 ? 0 : ((a0 >> 7) & 7);
            mma(tm, desc(a0 + k * 32, 16, 1280, bo), desc(s32(sb) + k * 32, 16, 1024, 0), idesc, k > 0);
        }
        commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float v[16];
    for (int c = 0; c < 64; c += 16) {
        ld16(tm + ((uint32_t)(warp * 32) << 16) + c, v);
        for (int i = 0; i < 16; ++i) out[(size_t)threadIdx.x * 64 + c + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}

// ---------------------------------------------------------------------------------------------- P2
// A: gy [32 pix][MB*32 ch] as MB tiles of [32 pix][32 ch] (TMA SW128), MN-major; B: x [32 pix][64 ch], 2 tiles, MN-major
// D[m][n] = sum_p A[p][m] * B[p][n].   variant 0: LBO = block stride, SBO = 1024 ; variant 1: swapped.
__global__ void probe2(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, int MB, int lbo, int sbo, int ltype, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = sm;                     // MB * 4096
    uint8_t* sb = sm + 4 * 4096;          // 2 * 4096
    uint64_t* bar = (uint64_t*)(sb + 2 * 4096);
    uint32_t* slot = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    if (threadIdx.x == 0) {
        mbar_expect(&bar[0], (MB + 2) * 4096);
        for (int b = 0; b < MB; ++b) tma2(&ma, &bar[0], sa + b * 4096, b * 32, 0);
        for (int b = 0; b < 2; ++b) tma2(&mb, &bar[0], sb + b * 4096, b * 32, 0);
        mbar_wait(&bar[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t M = MB * 32;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((M >> 4) << 24);
        for (int k = 0; k < 4; ++k) {
            mma(tm, desc(s32(sa) + k * 1024, lbo, sbo, 0, ltype), desc(s32(sb) + k * 1024, lbo, sbo, 0, ltype), idesc, k > 0);
        }
        commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float v[16];
    for (int c = 0; c < 64; c += 16) {
        ld16(tm + ((uint32_t)(warp * 32) << 16) + c, v);
        for (int i = 0; i < 16; ++i) out[(size_t)threadIdx.x * 64 + c + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}

static CUtensorMap make_map(void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                            CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
    CUtensorMap m;
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    float* out; CK(cudaMalloc(&out, 128 * 64 * 4));
    float* hout = (float*)malloc(128 * 64 * 4);
    // ---------------- P1
    {
        float* hh = (float*)malloc(18 * 10 * 32 * 4); float* hw = (float*)malloc(64 * 32 * 4);
        for (int i = 0; i < 18 * 10 * 32; ++i) hh[i] = (float)((i * 7 + (i / 32) * 3) % 17 - 8);
        for (int i = 0; i < 64 * 32; ++i) hw[i] = (float)((i * 5 + i / 32) % 9 - 4);
        float *dh, *dw; CK(cudaMalloc(&dh, 18 * 10 * 32 * 4)); CK(cudaMalloc(&dw, 64 * 32 * 4));
        CK(cudaMemcpy(dh, hh, 18 * 10 * 32 * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, hw, 64 * 32 * 4, cudaMemcpyHostToDevice));
        cuuint64_t d3[3] = {32, 10, 18}, s3[2] = {128, 1280}; cuuint32_t b3[3] = {32, 10, 18};
        cuuint64_t d2[2] = {32, 64}, s2[1] = {128}; cuuint32_t b2[2] = {32, 64};
        CUtensorMap mh = make_map(dh, 3, d3, s3, b3), mw = make_map(dw, 2, d2, s2, b2);
        CK(cudaFuncSetAttribute(probe1, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
        for (int mode = 0; mode < 2; ++mode)
            for (int t = 0; t < 5; ++t) {
                const int dys[5] = {0, 0, 1, 1, 2}, dxs[5] = {0, 1, 0, 1, 2};
                int dy = dys[t], dx = dxs[t];
                CK(cudaMemset(out, 0, 128 * 64 * 4));
                probe1<<<1, 128, 40000>>>(mh, mw, dy, dx, mode, out);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(hout, out, 128 * 64 * 4, cudaMemcpyDeviceToHost));
                double maxerr = 0; int bad = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < 64; ++n) {
                        int h = m / 8, w = m % 8; double ref = 0;
                        for (int c = 0; c < 32; ++c) ref += (double)hh[((h + dy) * 10 + (w + dx)) * 32 + c] * hw[n * 32 + c];
                        double e = fabs(ref - hout[m * 64 + n]); if (e > maxerr) maxerr = e; if (e > 1e-3) bad++;
                    }
                printf("P1 base_offset_mode=%d dy=%d dx=%d : maxerr %.3f bad %d/8192\n", mode, dy, dx, maxerr, bad);
            }
    }
    // ---------------- P2
    for (int MB = 2; MB <= 4; MB += 2) {
        const int M = MB * 32;
        float* ha = (float*)malloc(32 * M * 4); float* hb = (float*)malloc(32 * 64 * 4);
        for (int i = 0; i < 32 * M; ++i) ha[i] = (float)((i * 3 + i / M) % 11 - 5);
        for (int i = 0; i < 32 * 64; ++i) hb[i] = (float)((i * 7 + i / 64) % 13 - 6);
        float *da, *db; CK(cudaMalloc(&da, 32 * M * 4)); CK(cudaMalloc(&db, 32 * 64 * 4));
        CK(cudaMemcpy(da, ha, 32 * M * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb, 32 * 64 * 4, cudaMemcpyHostToDevice));
        cuuint64_t dA[2] = {(cuuint64_t)M, 32}, sA[1] = {(cuuint64_t)M * 4}; cuuint32_t bx[2] = {32, 32};
        cuuint64_t dB[2] = {64, 32}, sB[1] = {64 * 4};
        CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
        for (int swz = 0; swz < 2; ++swz) {
            CUtensorMapSwizzle sw = swz == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
            CUtensorMap ma = make_map(da, 2, dA, sA, bx, sw), mb = make_map(db, 2, dB, sB, bx, sw);
            const int lbos[6] = {4096, 1024, 4096, 512, 4096, 128}, sbos[6] = {1024, 4096, 512, 4096, 128, 4096};
            for (int variant = 0; variant < 6; ++variant) {
                const int ltype = swz == 0 ? 2 : 1;
                CK(cudaMemset(out, 0, 128 * 64 * 4));
                probe2<<<1, 128, 40000>>>(ma, mb, MB, lbos[variant], sbos[variant], ltype, out);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(hout, out, 128 * 64 * 4, cudaMemcpyDeviceToHost));
                for (int lm = 0; lm < 2; ++lm) {
                    double maxerr = 0; int bad = 0;
                    for (int m = 0; m < M; ++m)
                        for (int n = 0; n < 64; ++n) {
                            double ref = 0;
                            for (int p = 0; p < 32; ++p) ref += (double)ha[p * M + m] * hb[p * 64 + n];
                            int lane = lm == 0 ? m : (m / 16) * 32 + m % 16;
                            if (lane >= 128) { bad++; continue; }
                            double e = fabs(ref - hout[lane * 64 + n]); if (e > maxerr) maxerr = e; if (e > 1e-3) bad++;
                        }
                    printf("P2 M=%d tma_swizzle=%s ltype=%d LBO=%d SBO=%d lanemap=%d : maxerr %.3f bad %d/%d\n", M, swz ? "128B_ATOM_32B" : "128B", ltype,
                           lbos[variant], sbos[variant], lm, maxerr, bad, M * 64);
                }
            }
        }
    }
    printf("probe done\n");
    return 0;
}
