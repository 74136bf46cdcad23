// dynavsr_b200/csrc/update.cu
//
// Pixel losses and the fused MAML inner update.
//   * dvsr_loss_fwd: l1 / l2 / Charbonnier value AND gradient in one pass
//       (Video_base_model.py:39-50,191-195, loss.py:19-30, the 10*L1 SLR term test_dynavsr.py:274).
//   * dvsr_update_sgd / dvsr_update_adam: ONE launch over the flat EDVR+MFDN parameter buffer
//       (the reference runs torch.optim.{SGD,Adam}.step() per tensor over 158 tensors:
//        test_dynavsr.py:223-231,277; train_dynavsr.py:335-351,399).
#include "common.cuh"

namespace dvsr {

__global__ void loss_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ loss,
                            float* __restrict__ ga, long long n, int kind, float weight, float eps) {
    const float inv_n = 1.f / (float)n;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        float v, g;
        if (kind == DVSR_LOSS_L1) { v = fabsf(d); g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
        else if (kind == DVSR_LOSS_L2) { v = d * d; g = 2.f * d; }
        else if (kind == DVSR_LOSS_CB) { const float r = sqrtf(d * d + eps); v = r; g = d / r; }
        else {  // Huber with delta = eps (loss.py:5-17): 0.5 d^2 inside, delta |d| - 0.5 delta^2 outside
            const float ad = fabsf(d);
            if (ad <= eps) { v = 0.5f * d * d; g = d; }
            else { v = eps * ad - 0.5f * eps * eps; g = d > 0.f ? eps : -eps; }
        }
        acc += v;
        if (ga) ga[i] = weight * inv_n * g;
    }
    __shared__ float red[32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss, v * weight * inv_n);
    }
}

__global__ void scale_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ y, long long n) {
    const float k = __ldg(s);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = x[i] * k;
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, long long n, long long split, float lr0, float lr1,
                           float wd) {
    // torch.optim.SGD without momentum: g += wd * p (L2 weight decay); p -= lr * g
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float pi = p[i];
        p[i] = pi - (i < split ? lr0 : lr1) * (g[i] + wd * pi);
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, long long split, float lr0, float lr1, float b1, float b2, float eps,
                            float bc1, float bc2, float wd) {
    // torch.optim.Adam (single-tensor path): m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
    // p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
    const float rs = 1.f / sqrtf(bc2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float pi = p[i];
        const float gi = g[i] + wd * pi;         // torch.optim.Adam: L2 weight decay folded into the gradient
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float lr = i < split ? lr0 : lr1;
        p[i] = pi - (lr / bc1) * (mi / (sqrtf(vi) * rs + eps));
    }
}

// Exchange step of meta-training fused with the outer update (train_dynavsr.py:438 + the gradient averaging DDP would do):
// every rank reads the flat meta-gradient of ALL ranks straight from their HBM over NVLink (peer pointers of a symmetric-memory
// allocation), sums them in rank order (bit-identical on every rank), scales by 1/world and applies Adam / SGD to its copy of the
// meta-weights -- one pass, no reduced gradient buffer, no separate all-reduce launch.  Loads are cache-volatile: the peers
// rewrite these buffers every outer step.
template <bool ADAM>
__global__ void update_peers_kernel(float* __restrict__ p, const float* const* __restrict__ grads, int n_peers, long long goff, float scale,
                                    float* __restrict__ m, float* __restrict__ v, long long n, long long split, float lr0, float lr1,
                                    float b1, float b2, float eps, float bc1, float bc2, float wd) {
    const float rs = ADAM ? 1.f / sqrtf(bc2) : 0.f;
    const long long n4 = n >> 2;
    for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < n_peers; ++r) {
            const float4 t = __ldcv(reinterpret_cast<const float4*>(grads[r] + goff) + i4);
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
        float gv[4] = {g.x * scale, g.y * scale, g.z * scale, g.w * scale};
        float4 pv4 = reinterpret_cast<float4*>(p)[i4];
        float pv[4] = {pv4.x, pv4.y, pv4.z, pv4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long i = i4 * 4 + k;
            const float lr = i < split ? lr0 : lr1;
            const float gi = gv[k] + wd * pv[k];
            if (ADAM) {
                const float mi = b1 * m[i] + (1.f - b1) * gi;
                const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
                m[i] = mi;
                v[i] = vi;
                pv[k] = pv[k] - (lr / bc1) * (mi / (sqrtf(vi) * rs + eps));
            } else {
                pv[k] = pv[k] - lr * gi;
            }
        }
        reinterpret_cast<float4*>(p)[i4] = make_float4(pv[0], pv[1], pv[2], pv[3]);
    }
}

// The same exchange in reduce-scatter form, for more than a couple of ranks: the kernel above makes every rank read ALL peers'
// full gradients ((world - 1) x 84 MB for EDVR-L: 434 us at 4 GPUs, no better than NCCL).  Here rank r owns slice r of the flat
// buffer: it reads that slice from every rank (rank order -> one well-defined sum), applies Adam / SGD to its slice of the
// meta-weights (its moments for that slice only: the optimiser state is sharded), and writes the UPDATED WEIGHTS of the slice
// straight into the same slice of every rank's exchange buffer -- a region nobody else reads, because only the owner reduces
// it.  Per rank: (world - 1) / world of the buffer in over NVLink and the same amount out, in opposite directions at once, and
// 1 / world of the optimiser's HBM traffic.  After the closing barrier every rank's buffer holds the complete new meta-weights
// (bit-identical on all ranks by construction: each element was computed once), which the caller copies into place.
template <bool ADAM>
__global__ void update_peers_sliced_kernel(const float* __restrict__ p, float* const* __restrict__ bufs, int n_peers, int rank, long long goff,
                                           float scale, float* __restrict__ m, float* __restrict__ v, long long n, long long split,
                                           float lr0, float lr1, float b1, float b2, float eps, float bc1, float bc2, float wd) {
    const float rs = ADAM ? 1.f / sqrtf(bc2) : 0.f;
    const long long n4 = n >> 2;
    const long long per = (n4 + n_peers - 1) / n_peers;
    const long long lo = per * rank, hi = lo + per < n4 ? lo + per : n4;
    for (long long i4 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < hi; i4 += (long long)gridDim.x * blockDim.x) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < n_peers; ++r) {
            const float4 t = __ldcv(reinterpret_cast<const float4*>(bufs[r] + goff) + i4);
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
        float gv[4] = {g.x * scale, g.y * scale, g.z * scale, g.w * scale};
        const float4 pv4 = reinterpret_cast<const float4*>(p)[i4];
        float pv[4] = {pv4.x, pv4.y, pv4.z, pv4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long i = i4 * 4 + k;
            const float lr = i < split ? lr0 : lr1;
            const float gi = gv[k] + wd * pv[k];
            if (ADAM) {
                const float mi = b1 * m[i] + (1.f - b1) * gi;
                const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
                m[i] = mi;
                v[i] = vi;
                pv[k] = pv[k] - (lr / bc1) * (mi / (sqrtf(vi) * rs + eps));
            } else {
                pv[k] = pv[k] - lr * gi;
            }
        }
        const float4 out = make_float4(pv[0], pv[1], pv[2], pv[3]);
        for (int r = 0; r < n_peers; ++r) reinterpret_cast<float4*>(bufs[r] + goff)[i4] = out;
    }
    __threadfence_system();
}

__global__ void abs_sum_kernel(const float* __restrict__ x, float* __restrict__ out, long long npix, int pix_stride, int c0, int c1) {
    const int w = c1 - c0;
    const long long total = npix * w;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / w;
        const int c = (int)(i - p * w);
        acc += fabsf(__ldg(x + p * pix_stride + c0 + c));
    }
    __shared__ float red[32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(out, v);
    }
}

static int blocks_for(long long n) {
    long long b = (n + 255) / 256;
    if (b > sm_count() * 8) b = sm_count() * 8;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace dvsr

using namespace dvsr;
#define ST ((cudaStream_t)stream)

extern "C" int dvsr_loss_fwd(const float* a, const float* b, float* loss, float* ga, long long n, int kind, float weight,
                             float eps, void* stream) {
    DVSR_REQUIRE(a && b && loss && n > 0, "loss_fwd: bad arguments");
    DVSR_REQUIRE(kind >= DVSR_LOSS_L1 && kind <= DVSR_LOSS_HUBER, "loss_fwd: unknown loss kind %d", kind);
    loss_kernel<<<blocks_for(n), 256, 0, ST>>>(a, b, loss, ga, n, kind, weight, eps);
    return check_launch("loss_fwd");
}
extern "C" int dvsr_scale_by_device_scalar(const float* x, const float* s, float* y, long long n, void* stream) {
    DVSR_REQUIRE(x && s && y && n > 0, "scale_by_device_scalar: bad arguments");
    scale_kernel<<<blocks_for(n), 256, 0, ST>>>(x, s, y, n);
    return check_launch("scale_by_device_scalar");
}
extern "C" int dvsr_update_sgd(float* p, const float* g, long long n, long long split, float lr0, float lr1, float wd,
                               void* stream) {
    DVSR_REQUIRE(p && g && n > 0, "update_sgd: bad arguments");
    sgd_kernel<<<blocks_for(n), 256, 0, ST>>>(p, g, n, split, lr0, lr1, wd);
    return check_launch("update_sgd");
}
extern "C" int dvsr_update_adam(float* p, const float* g, float* m, float* v, long long n, long long split, float lr0,
                                float lr1, float b1, float b2, float eps, float bc1, float bc2, float wd, void* stream) {
    DVSR_REQUIRE(p && g && m && v && n > 0, "update_adam: bad arguments");
    adam_kernel<<<blocks_for(n), 256, 0, ST>>>(p, g, m, v, n, split, lr0, lr1, b1, b2, eps, bc1, bc2, wd);
    return check_launch("update_adam");
}
extern "C" int dvsr_update_peers(float* p, const float* const* grads_dev, int n_peers, long long grad_offset, float scale, float* m,
                                 float* v, long long n, long long split, float lr0, float lr1, float b1, float b2, float eps,
                                 float bc1, float bc2, float wd, int adam, void* stream) {
    DVSR_REQUIRE(p && grads_dev && n_peers >= 1 && n > 0 && (n & 3) == 0, "update_peers: bad arguments (n must be a multiple of 4)");
    DVSR_REQUIRE(!adam || (m && v), "update_peers: Adam needs the moment buffers");
    DVSR_REQUIRE((((uintptr_t)p) & 15) == 0 && (grad_offset & 3) == 0, "update_peers: 16-byte alignment required");
    if (adam) update_peers_kernel<true><<<blocks_for(n >> 2), 256, 0, ST>>>(p, grads_dev, n_peers, grad_offset, scale, m, v, n, split, lr0, lr1, b1, b2, eps, bc1, bc2, wd);
    else update_peers_kernel<false><<<blocks_for(n >> 2), 256, 0, ST>>>(p, grads_dev, n_peers, grad_offset, scale, m, v, n, split, lr0, lr1, b1, b2, eps, bc1, bc2, wd);
    return check_launch("update_peers");
}
extern "C" int dvsr_update_peers_sliced(const float* p, float* const* bufs_dev, int n_peers, int rank, long long buf_offset, float scale,
                                        float* m, float* v, long long n, long long split, float lr0, float lr1, float b1, float b2,
                                        float eps, float bc1, float bc2, float wd, int adam, void* stream) {
    DVSR_REQUIRE(p && bufs_dev && n_peers >= 1 && rank >= 0 && rank < n_peers && n > 0 && (n & 3) == 0,
                 "update_peers_sliced: bad arguments (n must be a multiple of 4)");
    DVSR_REQUIRE(!adam || (m && v), "update_peers_sliced: Adam needs the moment buffers");
    DVSR_REQUIRE((((uintptr_t)p) & 15) == 0 && (buf_offset & 3) == 0, "update_peers_sliced: 16-byte alignment required");
    const long long per = ((n >> 2) + n_peers - 1) / n_peers;
    if (adam) update_peers_sliced_kernel<true><<<blocks_for(per), 256, 0, ST>>>(p, bufs_dev, n_peers, rank, buf_offset, scale, m, v, n, split, lr0, lr1, b1, b2, eps, bc1, bc2, wd);
    else update_peers_sliced_kernel<false><<<blocks_for(per), 256, 0, ST>>>(p, bufs_dev, n_peers, rank, buf_offset, scale, m, v, n, split, lr0, lr1, b1, b2, eps, bc1, bc2, wd);
    return check_launch("update_peers_sliced");
}
extern "C" int dvsr_abs_sum(const float* x, float* out, long long npix, int pix_stride, int c0, int c1, void* stream) {
    DVSR_REQUIRE(x && out && npix > 0 && c1 > c0 && pix_stride >= c1, "abs_sum: bad arguments");
    abs_sum_kernel<<<blocks_for(npix * (c1 - c0)), 256, 0, ST>>>(x, out, npix, pix_stride, c0, c1);
    return check_launch("abs_sum");
}
