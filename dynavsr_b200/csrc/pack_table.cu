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

__global__ void pack_table_kernel(const dvsr_pack_job* __restrict__ table, int n) {
    // binary search: last job whose block_start <= blockIdx.x
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].block_start <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const dvsr_pack_job j = table[lo];
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
