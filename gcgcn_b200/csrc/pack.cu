// Parameter packing between the reference's per-GraphConv parameters (weights_node [128 + l*g, g],
// weights_edge [128, g], G:24-25) and the column-concatenated layouts the batched kernels consume
// (include/gcgcn_b200.h): one launch each way instead of ~40 small copy kernels per step.
#include <cmath>

#include "common.cuh"

namespace gcgcn {

// grid.x = heads*layers (one GraphConv each), grid.y = slices
__global__ void __launch_bounds__(256)
pack_stack_kernel(const float* const* __restrict__ wn_ptrs, const float* const* __restrict__ we_ptrs,
                  int heads, int layers, int slab, float* __restrict__ WnX, float* __restrict__ We,
                  float* __restrict__ Winner) {
    const int k = blockIdx.x, h = k / layers, l = k - h * layers;
    const int g = slab / layers, HD = heads * slab;
    const float* wn = wn_ptrs[k];
    const float* we = we_ptrs[k];
    const int rows_n = D + l * g;
    const int stride = gridDim.y * blockDim.x;
    const int t0 = blockIdx.y * blockDim.x + threadIdx.x;
    for (int idx = t0; idx < rows_n * g; idx += stride) {
        const int r = idx / g, c = idx - r * g;
        const float v = wn[idx];
        if (r < D) WnX[static_cast<size_t>(r) * HD + k * g + c] = v;
        else if (Winner != nullptr) Winner[(static_cast<size_t>(k) * slab + (r - D)) * g + c] = v;
    }
    if (Winner != nullptr)
        for (int idx = t0 + l * g * g; idx < slab * g; idx += stride)      // unused rows of the block
            Winner[static_cast<size_t>(k) * slab * g + idx] = 0.f;
    for (int idx = t0; idx < D * g; idx += stride) {
        const int r = idx / g, c = idx - r * g;
        We[static_cast<size_t>(r) * HD + k * g + c] = we[idx];
    }
}

// inverse map for the gradients; outputs are two flat buffers holding the per-GraphConv gradients back to back
__global__ void __launch_bounds__(256)
unpack_stack_kernel(const float* __restrict__ dWnX, const float* __restrict__ dWe,
                    const float* __restrict__ dWinner, int heads, int layers, int slab,
                    float* __restrict__ dwn_flat, float* __restrict__ dwe_flat) {
    const int k = blockIdx.x, h = k / layers, l = k - h * layers;
    const int g = slab / layers, HD = heads * slab;
    const int rows_n = D + l * g;
    // offset of GraphConv k inside dwn_flat: h full heads + l earlier sub-layers of this head
    const size_t per_head = static_cast<size_t>(layers) * D * g + static_cast<size_t>(g) * g * (layers * (layers - 1) / 2);
    const size_t off_n = h * per_head + static_cast<size_t>(l) * D * g + static_cast<size_t>(g) * g * (l * (l - 1) / 2);
    float* dwn = dwn_flat + off_n;
    float* dwe = dwe_flat + static_cast<size_t>(k) * D * g;
    const int stride = gridDim.y * blockDim.x;
    const int t0 = blockIdx.y * blockDim.x + threadIdx.x;
    for (int idx = t0; idx < rows_n * g; idx += stride) {
        const int r = idx / g, c = idx - r * g;
        dwn[idx] = (r < D) ? dWnX[static_cast<size_t>(r) * HD + k * g + c]
                           : dWinner[(static_cast<size_t>(k) * slab + (r - D)) * g + c];
    }
    for (int idx = t0; idx < D * g; idx += stride) {
        const int r = idx / g, c = idx - r * g;
        dwe[idx] = dWe[static_cast<size_t>(r) * HD + k * g + c];
    }
}

int launch_pack_stack(const float* const* wn_ptrs, const float* const* we_ptrs, int heads, int layers, int slab,
                      float* WnX, float* We, float* Winner, cudaStream_t st) {
    dim3 grid(heads * layers, 8);
    pack_stack_kernel<<<grid, 256, 0, st>>>(wn_ptrs, we_ptrs, heads, layers, slab, WnX, We, Winner);
    GCGCN_CHECK_LAUNCH("pack_stack_weights");
    return GCGCN_OK;
}

int launch_unpack_stack(const float* dWnX, const float* dWe, const float* dWinner, int heads, int layers, int slab,
                        float* dwn_flat, float* dwe_flat, cudaStream_t st) {
    dim3 grid(heads * layers, 8);
    unpack_stack_kernel<<<grid, 256, 0, st>>>(dWnX, dWe, dWinner, heads, layers, slab, dwn_flat, dwe_flat);
    GCGCN_CHECK_LAUNCH("unpack_stack_grads");
    return GCGCN_OK;
}

// ---- fused Adam over the flat parameter bucket (config 5: one optimiser step per micro-batch) ----
// torch.optim.Adam semantics (the reference trains with optim.Adam(lr), C:300): decoupled from autograd,
// one pass over four flat fp32 arrays, float4 accesses.  grad_scale folds the 1/world (or 1/documents)
// averaging of the all-reduced gradient sum into the same pass.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long count, float lr, float b1, float b2, float eps, float wd, float gscale, float inv_bias1,
            float inv_sqrt_bias2) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long t0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg = gg * gscale + wd * pp;
        mm = mm + (gg - mm) * (1.f - b1);
        vv = b2 * vv + (1.f - b2) * gg * gg;
        const float denom = sqrtf(vv) * inv_sqrt_bias2 + eps;
        pp -= lr * inv_bias1 * (mm / denom);
    };
    const long long n4 = count >> 2;
    for (long long i = t0; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i],
               vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = (n4 << 2) + t0; i < count; i += stride) upd(p[i], g[i], m[i], v[i]);
}

int launch_adam(float* p, const float* g, float* m, float* v, long long count, float lr, float b1, float b2,
                float eps, float wd, float gscale, int step, cudaStream_t st) {
    const double bias1 = 1.0 - pow(static_cast<double>(b1), step), bias2 = 1.0 - pow(static_cast<double>(b2), step);
    long long blocks = (count / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(p, g, m, v, count, lr, b1, b2, eps, wd, gscale,
                                                             static_cast<float>(1.0 / bias1),
                                                             static_cast<float>(1.0 / sqrt(bias2)));
    GCGCN_CHECK_LAUNCH("adam_step");
    return GCGCN_OK;
}

}  // namespace gcgcn
