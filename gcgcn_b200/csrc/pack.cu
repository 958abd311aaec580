// Parameter packing between the reference's per-GraphConv parameters (weights_node [128 + l*g, g],
// weights_edge [128, g], G:24-25) and the column-concatenated layouts the batched kernels consume
// (include/gcgcn_b200.h): one launch each way instead of ~40 small copy kernels per step.
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

}  // namespace gcgcn
