// dynavsr_b200/csrc/pack_table.cu -- (re)packing of convolution weights into kernel layouts.
//   dvsr_pack_job   : one weight, one layout, one launch (used lazily at first use);
//   dvsr_pack_table : a device-resident table of jobs, ONE launch for the whole model -- issued after every
//                     fused parameter update / restore instead of ~370 tiny per-layer launches.
#include "pack_device.cuh"

namespace dvsr {

__global__ void pack_job_kernel(const dvsr_pack_job j) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < j.total) j.wp[i] = pack_value(j, i);
}

// BF16x3 stacked layout (modes 9 / 10, the bulk of a model's packed bytes): the 256 elements of one CUDA block are 8
// consecutive rows of ONE 128-row weight block, so (output group, segment, channel pair, tap) are block-uniform -- the
// runtime divisions of pack_value are done once per block instead of once per element.
__device__ __forceinline__ void pack_block_stacked(const dvsr_pack_job& j, long long blk_in_job) {
    __shared__ int sh[6];      // n0 (first output row), lo, seg, pair, tap, valid
    const dvsr_wlayout& wl = j.wl;
    if (threadIdx.x == 0) {
        long long r = blk_in_job * 8;                 // first of the 8 rows
        const int row = (int)(r % 128);
        r /= 128;
        int blk = (int)(r % j.a0);
        const int g = (int)(r / j.a0);
        int s = j.seg, pair, tap;
        if (j.mode == 9) {
            for (; s < j.seg_hi; ++s) {
                const int nb = wl.taps * ((wl.seg_C[s] + 63) / 64);
                if (blk < nb) break;
                blk -= nb;
            }
        }
        pair = blk / wl.taps;
        tap = blk - pair * wl.taps;
        sh[0] = g * 64 + (row & 63); sh[1] = row >= 64; sh[2] = s; sh[3] = pair; sh[4] = tap;
    }
    __syncthreads();
    const int k2 = threadIdx.x & 31;
    const int n = sh[0] + (threadIdx.x >> 5);
    const bool want_lo = sh[1] != 0;
    const int s = sh[2], pair = sh[3], tap = sh[4];
    const float* __restrict__ w = j.w;
    uint32_t out = 0;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int ch = pair * 64 + 2 * k2 + e;
        float v = 0.f;
        if (j.mode == 9) {
            const long long co_ = (n < wl.Co && ch < wl.seg_C[s]) ? wl_ci_offset(wl, ch) : -1;
            if (co_ >= 0) v = __ldg(w + (long long)n * wl.co_stride + wl.seg_base[s] + co_ + tap);
        } else if (n < wl.seg_C[j.seg] && ch < wl.Co) {
            v = __ldg(w + (long long)ch * wl.co_stride + wl.seg_base[j.seg] + (long long)n * wl.ci_stride + tap);
        }
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        out |= (uint32_t)__bfloat16_as_ushort(want_lo ? l : h) << (16 * e);
    }
    j.wp[blk_in_job * 256 + threadIdx.x] = __uint_as_float(out);
}

__global__ void pack_table_kernel(const dvsr_pack_job* __restrict__ table, int n) {
    // binary search: last job whose block_start <= blockIdx.x
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].block_start <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const dvsr_pack_job& j = table[lo];
    if (j.mode >= 9) { pack_block_stacked(j, (long long)blockIdx.x - j.block_start); return; }   // total % 256 == 0
    const long long i = ((long long)blockIdx.x - j.block_start) * blockDim.x + threadIdx.x;
    if (i < j.total) j.wp[i] = pack_value(j, i);
}

// Snapshot / restore of every packed buffer of a table to / from ONE arena (arena offset of job j = block_start * 256
// floats): "deepcopy per frame" (test_dynavsr.py:208) restores the packs of the meta-weights with a plain copy instead
// of re-deriving them from the restored weights.
__global__ void pack_table_copy_kernel(const dvsr_pack_job* __restrict__ table, int n, float* __restrict__ arena, int to_packs) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].block_start <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const dvsr_pack_job& j = table[lo];
    const long long i = ((long long)blockIdx.x - j.block_start) * blockDim.x + threadIdx.x;
    if (i >= j.total) return;
    float* a = arena + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (to_packs) j.wp[i] = *a; else *a = j.wp[i];
}

}  // namespace dvsr

using namespace dvsr;

extern "C" int dvsr_pack_table_copy(const dvsr_pack_job* table_dev, int n_jobs, long long total_blocks, float* arena,
                                    int to_packs, void* stream) {
    DVSR_REQUIRE(table_dev && arena && n_jobs > 0 && total_blocks > 0 && total_blocks < 0x7fffffffLL, "pack_table_copy: bad arguments");
    pack_table_copy_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(table_dev, n_jobs, arena, to_packs);
    return check_launch("pack_table_copy");
}

extern "C" int dvsr_pack_job_run(const dvsr_pack_job* job, void* stream) {
    DVSR_REQUIRE(job && job->w && job->wp && job->total > 0 && job->mode >= 0 && job->mode <= 10, "pack_job: bad job");
    pack_job_kernel<<<cdiv(job->total, 256), 256, 0, (cudaStream_t)stream>>>(*job);
    return check_launch("pack_job");
}

extern "C" int dvsr_pack_table(const dvsr_pack_job* table_dev, int n_jobs, long long total_blocks, void* stream) {
    DVSR_REQUIRE(table_dev && n_jobs > 0 && total_blocks > 0 && total_blocks < 0x7fffffffLL, "pack_table: bad arguments");
    pack_table_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(table_dev, n_jobs);
    return check_launch("pack_table");
}
