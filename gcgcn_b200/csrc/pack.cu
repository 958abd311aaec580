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

// ---- GATAttention parameter collapse (G:156-162) ------------------------------------------------------------
// energy_ij = wt . [Wh x_j + bh ; Wt x_j + bt ; Wr e_ij + br] + b  ==  u . x_j + v . e_ij + c  with
//   u = Wh^T w1 + Wt^T w2,  v = Wr^T w3,  c = w1.bh + w2.bt + w3.br + b        (w = wt.weight = [w1 | w2 | w3])
// One CTA of 8 row groups x 128 columns: thread (g, k) sums rows g, g + 8, ... of column k of the three [hid, 128]
// weights (coalesced rows, independent loads in flight; one thread per column walked 128 dependent rows: 54 us),
// partial sums meet in shared memory.  out = [u(128) | v(128) | c].
constexpr int GC_GROUPS = 8;
__global__ void __launch_bounds__(GC_GROUPS * D)
gat_collapse_fwd_kernel(const float* __restrict__ Wh, const float* __restrict__ bh, const float* __restrict__ Wt,
                        const float* __restrict__ bt, const float* __restrict__ Wr, const float* __restrict__ br,
                        const float* __restrict__ w, const float* __restrict__ b, int hid, float* __restrict__ out) {
    __shared__ float su[GC_GROUPS][D], sv[GC_GROUPS][D], red[GC_GROUPS * D / 32];
    const int k = threadIdx.x % D, g = threadIdx.x / D;
    float u = 0.f, v = 0.f;
#pragma unroll 4
    for (int o = g; o < hid; o += GC_GROUPS) {
        u += Wh[o * D + k] * w[o] + Wt[o * D + k] * w[hid + o];
        v += Wr[o * D + k] * w[2 * hid + o];
    }
    su[g][k] = u;
    sv[g][k] = v;
    float c = 0.f;
    for (int o = threadIdx.x; o < hid; o += GC_GROUPS * D) c += w[o] * bh[o] + w[hid + o] * bt[o] + w[2 * hid + o] * br[o];
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (g == 0) {
        float tu = 0.f, tv = 0.f;
#pragma unroll
        for (int i = 0; i < GC_GROUPS; ++i) { tu += su[i][k]; tv += sv[i][k]; }
        out[k] = tu;
        out[D + k] = tv;
        if (k == 0) {
            float tc = b[0];
#pragma unroll
            for (int i = 0; i < GC_GROUPS * D / 32; ++i) tc += red[i];
            out[2 * D] = tc;
        }
    }
}

// gradients of all eight parameters from (du, dv, dc); grid = hid rows, thread k = column
__global__ void __launch_bounds__(128)
gat_collapse_bwd_kernel(const float* __restrict__ Wh, const float* __restrict__ bh, const float* __restrict__ Wt,
                        const float* __restrict__ bt, const float* __restrict__ Wr, const float* __restrict__ br,
                        const float* __restrict__ w, const float* __restrict__ dout, int hid, float* __restrict__ dWh,
                        float* __restrict__ dbh, float* __restrict__ dWt, float* __restrict__ dbt, float* __restrict__ dWr,
                        float* __restrict__ dbr, float* __restrict__ dw, float* __restrict__ db) {
    const int o = blockIdx.x, k = threadIdx.x;
    const float du = dout[k], dv = dout[D + k], dc = dout[2 * D];
    const float w1 = w[o], w2 = w[hid + o], w3 = w[2 * hid + o];
    dWh[o * D + k] = w1 * du;
    dWt[o * D + k] = w2 * du;
    dWr[o * D + k] = w3 * dv;
    float a1 = Wh[o * D + k] * du, a2 = Wt[o * D + k] * du, a3 = Wr[o * D + k] * dv;
    a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    __shared__ float red[3][4];
    if ((k & 31) == 0) { red[0][k >> 5] = a1; red[1][k >> 5] = a2; red[2][k >> 5] = a3; }
    __syncthreads();
    if (k == 0) {
        dw[o] = red[0][0] + red[0][1] + red[0][2] + red[0][3] + dc * bh[o];
        dw[hid + o] = red[1][0] + red[1][1] + red[1][2] + red[1][3] + dc * bt[o];
        dw[2 * hid + o] = red[2][0] + red[2][1] + red[2][2] + red[2][3] + dc * br[o];
        dbh[o] = dc * w1;
        dbt[o] = dc * w2;
        dbr[o] = dc * w3;
        if (o == 0) db[0] = dc;
    }
}

int launch_gat_collapse_fwd(const float* Wh, const float* bh, const float* Wt, const float* bt, const float* Wr,
                            const float* br, const float* w, const float* b, int hid, float* out, cudaStream_t st) {
    gat_collapse_fwd_kernel<<<1, GC_GROUPS * D, 0, st>>>(Wh, bh, Wt, bt, Wr, br, w, b, hid, out);
    GCGCN_CHECK_LAUNCH("gat_collapse_fwd");
    return GCGCN_OK;
}
int launch_gat_collapse_bwd(const float* Wh, const float* bh, const float* Wt, const float* bt, const float* Wr,
                            const float* br, const float* w, const float* dout, int hid, float* dWh, float* dbh, float* dWt,
                            float* dbt, float* dWr, float* dbr, float* dw, float* db, cudaStream_t st) {
    gat_collapse_bwd_kernel<<<hid, 128, 0, st>>>(Wh, bh, Wt, bt, Wr, br, w, dout, hid, dWh, dbh, dWt, dbt, dWr, dbr, dw, db);
    GCGCN_CHECK_LAUNCH("gat_collapse_bwd");
    return GCGCN_OK;
}

// ---- row-wise concatenation of equally shaped matrices through a pointer table (MultiHeadAttention's H query
// projections [d_h, 128] + biases -> Wq [128, 128], bq [128]: one launch instead of two torch.cat) -------------------
__global__ void __launch_bounds__(256)
pack_rows_kernel(const float* const* __restrict__ ptrs, int count, int elems, float* __restrict__ out) {
    const int i = blockIdx.y;
    const float* src = ptrs[i];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < elems; idx += gridDim.x * blockDim.x)
        out[static_cast<size_t>(i) * elems + idx] = src[idx];
}
int launch_pack_rows(const float* const* ptrs, int count, int elems, float* out, cudaStream_t st) {
    if (count <= 0 || elems <= 0) return GCGCN_OK;
    dim3 grid(std::max(1, std::min(8, (elems + 255) / 256)), count);
    pack_rows_kernel<<<grid, 256, 0, st>>>(ptrs, count, elems, out);
    GCGCN_CHECK_LAUNCH("pack_rows");
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
