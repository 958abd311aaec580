// Mention->entity pooling (G:297-298) and head/tail pair gathers (G:351-352, 321-322) as
// index-driven, 128-bit vectorised row gathers instead of dense mapping-matrix products.
#include "common.cuh"

namespace gcgcn {

int launch_reduce_partials(const float* partial, int parts, int width, float* out0, int width0,
                           float* out1, cudaStream_t st);

// out[row,:] = sum_{k in [ptr[row], ptr[row+1])} w[k] * src[idx[k],:]      (one warp per row)
// forward: rows = entities, idx = tokens; backward: rows = tokens, idx = entities.
__global__ void __launch_bounds__(256)
csr_gather_kernel(const float* __restrict__ src, const int* __restrict__ ptr,
                  const int* __restrict__ idx, const float* __restrict__ w, int rows,
                  float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < rows; r += nwarps) {
        const int k0 = ptr[r], k1 = ptr[r + 1];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = k0; k < k1; ++k) {
            const float wk = w[k];
            const float4 v = *reinterpret_cast<const float4*>(src + static_cast<size_t>(idx[k]) * D + lane * 4);
            acc.x = fmaf(wk, v.x, acc.x); acc.y = fmaf(wk, v.y, acc.y);
            acc.z = fmaf(wk, v.z, acc.z); acc.w = fmaf(wk, v.w, acc.w);
        }
        *reinterpret_cast<float4*>(out + static_cast<size_t>(r) * D + lane * 4) = acc;
    }
}

int launch_csr_gather(const float* src, const int* ptr, const int* idx, const float* w, int rows,
                      float* out, cudaStream_t st) {
    if (rows == 0) return GCGCN_OK;
    const int blocks = min(ceil_div(rows, 8), sm_count() * 8);
    csr_gather_kernel<<<blocks, 256, 0, st>>>(src, ptr, idx, w, rows, out);
    GCGCN_CHECK_LAUNCH("csr_gather");
    return GCGCN_OK;
}

// ---------------------------------------------------------------------------------------------
// pair gather forward: one warp per pair writes both output rows (h and t side) of feat_w + dis_w floats.
// All index loads of a pair are issued together, then all row loads (CH float4 per lane and side), then the
// stores: 2*CH independent 16-byte loads in flight per lane instead of a idx -> row -> store chain per row.
template <int CH>
__global__ void __launch_bounds__(256)
pair_gather_fwd_kernel(const float4* __restrict__ feat, int fw4, const float4* __restrict__ dis, int dw4,
                       const int* __restrict__ h_idx, const int* __restrict__ t_idx,
                       const int* __restrict__ dis_h, const int* __restrict__ dis_t,
                       float* __restrict__ out_h, float* __restrict__ out_t, long long total_pairs) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int ow4 = fw4 + dw4;
    for (long long p = warp; p < total_pairs; p += nwarps) {
        const float4* fh = feat + static_cast<size_t>(h_idx[p]) * fw4;
        const float4* ft = feat + static_cast<size_t>(t_idx[p]) * fw4;
        const float4* dh = dw4 > 0 ? dis + static_cast<size_t>(dis_h[p]) * dw4 : nullptr;
        const float4* dt = dw4 > 0 ? dis + static_cast<size_t>(dis_t[p]) * dw4 : nullptr;
        float* oh = out_h + static_cast<size_t>(p) * ow4 * 4;
        float* ot = out_t + static_cast<size_t>(p) * ow4 * 4;
        for (int q0 = 0; q0 < ow4; q0 += CH * WARP) {
            float4 vh[CH], vt[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int q = q0 + c * WARP + lane;
                if (q < ow4) {
                    vh[c] = q < fw4 ? __ldg(fh + q) : __ldg(dh + (q - fw4));
                    vt[c] = q < fw4 ? __ldg(ft + q) : __ldg(dt + (q - fw4));
                }
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int q = q0 + c * WARP + lane;
                if (q < ow4) {
                    Vec4<float>::store(oh + q * 4, vh[c]);
                    Vec4<float>::store(ot + q * 4, vt[c]);
                }
            }
        }
    }
}

// pair gather backward, feature part: node row (b,k) was gathered by the "h" side of every pair
// (i,k) and by the "t" side of every pair (k,j)  (G:351-352: h gathers column j, t gathers row i).
__global__ void __launch_bounds__(128)
pair_gather_bwd_feat_kernel(const int* __restrict__ node_ptr, const long long* __restrict__ pair_ptr,
                            const int* __restrict__ row_doc, const float* __restrict__ dout_h,
                            const float* __restrict__ dout_t, int fw4, int ow4,
                            float* __restrict__ dfeat) {
    const int r = blockIdx.x;
    const int b = row_doc[r];
    const int node0 = node_ptr[b];
    const int n = node_ptr[b + 1] - node0;
    const int k = r - node0;
    const long long p0 = pair_ptr[b];
    for (int q = threadIdx.x; q < fw4; q += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const float4 v = Vec4<float>::load(dout_h + (static_cast<size_t>(p0 + static_cast<long long>(i) * n + k) * ow4 + q) * 4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float4 v = Vec4<float>::load(dout_t + (static_cast<size_t>(p0 + static_cast<long long>(k) * n + j) * ow4 + q) * 4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(dfeat + (static_cast<size_t>(r) * fw4 + q) * 4) = acc;
    }
}

// pair gather backward, distance-table part: every warp owns a contiguous chunk of pairs and a
// private [dis_rows][dis_w] accumulator in shared memory (lane = column), walked sequentially so
// the summation order is fixed; per-warp partials are reduced by reduce_partials.
constexpr int DIS_BWD_THREADS = 128;
__global__ void __launch_bounds__(DIS_BWD_THREADS)
pair_gather_bwd_dis_kernel(const float* __restrict__ dout_h, const float* __restrict__ dout_t, int fw,
                           int dw, int dis_rows, const int* __restrict__ dis_h,
                           const int* __restrict__ dis_t, long long total_pairs,
                           long long pairs_per_warp, float* __restrict__ partial) {
    extern __shared__ float acc_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cells = dis_rows * dw;
    float* acc = acc_all + warp * cells;
    for (int t = lane; t < cells; t += WARP) acc[t] = 0.f;
    __syncwarp();
    const long long gw = static_cast<long long>(blockIdx.x) * (DIS_BWD_THREADS / WARP) + warp;
    const long long p0 = gw * pairs_per_warp;
    const long long p1 = min(total_pairs, p0 + pairs_per_warp);
    const int ow = fw + dw;
    // eight pairs per trip: their index and gradient loads are issued together, the shared-memory updates then
    // run in pair order (fixed summation order); the chain per pair was global load -> shared read-modify-write
    constexpr int U = 8;
    for (long long p = p0; p < p1; p += U) {
        int kh[U], kt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long pp = min(p + u, p1 - 1);
            kh[u] = dis_h[pp];
            kt[u] = dis_t[pp];
        }
        for (int c = lane; c < dw; c += WARP) {
            float vh[U], vt[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long pp = min(p + u, p1 - 1);
                vh[u] = __ldg(dout_h + static_cast<size_t>(pp) * ow + fw + c);
                vt[u] = __ldg(dout_t + static_cast<size_t>(pp) * ow + fw + c);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (p + u < p1) {
                    acc[kh[u] * dw + c] += vh[u];
                    acc[kt[u] * dw + c] += vt[u];
                }
            }
        }
    }
    __syncwarp();
    for (int t = lane; t < cells; t += WARP) partial[static_cast<size_t>(gw) * cells + t] = acc[t];
}

int launch_pair_gather_fwd(const gcgcn_batch* bt, const float* feat, int feat_w, const float* dis,
                           int dis_w, const int* h_idx, const int* t_idx, const int* dis_h,
                           const int* dis_t, float* out_h, float* out_t, cudaStream_t st) {
    if (bt->total_pairs == 0) return GCGCN_OK;
    const long long warps = bt->total_pairs;
    const int blocks = static_cast<int>(std::min<long long>((warps + 7) / 8, static_cast<long long>(sm_count()) * 8));
    const int ow4 = (feat_w + dis_w) / 4;
    auto kern = ow4 <= WARP ? pair_gather_fwd_kernel<1> : ow4 <= 2 * WARP ? pair_gather_fwd_kernel<2>
                                                                         : pair_gather_fwd_kernel<4>;
    kern<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(feat), feat_w / 4,
                                 reinterpret_cast<const float4*>(dis), dis_w / 4, h_idx, t_idx, dis_h, dis_t,
                                 out_h, out_t, bt->total_pairs);
    GCGCN_CHECK_LAUNCH("pair_gather_fwd");
    return GCGCN_OK;
}

int pair_dis_warps() { return sm_count() * 16; }

int launch_pair_gather_bwd(const gcgcn_batch* bt, const float* dout_h, const float* dout_t, int feat_w,
                           int dis_w, int dis_rows, const int* dis_h, const int* dis_t, float* dfeat,
                           float* ddis, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (bt->total_nodes == 0) return GCGCN_OK;
    const int ow4 = (feat_w + dis_w) / 4;
    pair_gather_bwd_feat_kernel<<<bt->total_nodes, 128, 0, st>>>(
        bt->node_ptr, reinterpret_cast<const long long*>(bt->pair_ptr), bt->row_doc, dout_h, dout_t,
        feat_w / 4, ow4, dfeat);
    GCGCN_CHECK_LAUNCH("pair_gather_bwd_feat");
    if (dis_w > 0 && ddis != nullptr) {
        const int cells = dis_rows * dis_w;
        long long warps = std::min<long long>(pair_dis_warps(), std::max<long long>(1, bt->total_pairs / 64));
        const int wpb = DIS_BWD_THREADS / WARP;
        const int blocks = ceil_div(warps, wpb);
        warps = static_cast<long long>(blocks) * wpb;
        const long long ppw = (bt->total_pairs + warps - 1) / warps;
        const size_t need = static_cast<size_t>(warps) * cells * sizeof(float);
        if (ws == nullptr || ws_bytes < need)
            return fail(GCGCN_ERR_WORKSPACE, "pair_gather_bwd: workspace %zu < %zu bytes", ws_bytes, need);
        const size_t smem = static_cast<size_t>(wpb) * cells * sizeof(float);
        if (smem > 48 * 1024)
            return fail(GCGCN_ERR_UNSUPPORTED, "pair_gather_bwd: distance table %d x %d too large", dis_rows, dis_w);
        pair_gather_bwd_dis_kernel<<<blocks, DIS_BWD_THREADS, smem, st>>>(
            dout_h, dout_t, feat_w, dis_w, dis_rows, dis_h, dis_t, bt->total_pairs, ppw,
            static_cast<float*>(ws));
        GCGCN_CHECK_LAUNCH("pair_gather_bwd_dis");
        GCGCN_TRY(launch_reduce_partials(static_cast<const float*>(ws), static_cast<int>(warps), cells,
                                         ddis, cells, nullptr, st));
    }
    return GCGCN_OK;
}

// ---------------------------------------------------------------------------------------------
// Classifier-side pair features without the gathered intermediate (first half):
//   entity_feature_h[i,j] = tanh(dense_layer(cat(F[j], dis[10 + rp_ij])))  (G:351-355)
//                         = tanh(U[j] + Vd[10 + rp_ij]),   U = F W_F^T  [rows, 128],  Vd = dis W_d^T + b  [21, 128]
// (dense_layer's weight split by input columns), and likewise entity_feature_t[i,j] = tanh(U[i] + Vd[10 - rp_ij]).
// The two n^2 x 424 gathered tensors and the n^2 x 424 x 128 product over them are never formed: one warp per pair
// reads two 512-byte rows per side (L2-resident tables) and writes 2 x 128 floats.
// tanh(x) = 1 - 2 / (exp(2x) + 1): ~6 instructions instead of tanhf's ~25, absolute error ~1e-7 (the parity bar is an
// absolute 1e-4); saturates correctly (exp -> inf gives 1, exp -> 0 gives -1)
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

__global__ void __launch_bounds__(256)
pair_dense_fwd_kernel(const float4* __restrict__ U, const float4* __restrict__ Vd, const int* __restrict__ h_idx,
                      const int* __restrict__ t_idx, const int* __restrict__ dis_h, const int* __restrict__ dis_t,
                      float* __restrict__ out_h, float* __restrict__ out_t, long long total_pairs) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    // two pairs per trip: eight index loads, then eight 512-byte row reads in flight, then four row stores
    for (long long p = warp; p < total_pairs; p += 2 * nwarps) {
        const long long p2 = p + nwarps;
        const bool two = p2 < total_pairs;
        const long long q = two ? p2 : p;
        const int ih = h_idx[p], it = t_idx[p], kh = dis_h[p], kt = dis_t[p];
        const int jh = h_idx[q], jt = t_idx[q], lh = dis_h[q], lt = dis_t[q];
        const float4 uh = __ldg(U + static_cast<size_t>(ih) * (D / 4) + lane);
        const float4 ut = __ldg(U + static_cast<size_t>(it) * (D / 4) + lane);
        const float4 vh = __ldg(Vd + static_cast<size_t>(kh) * (D / 4) + lane);
        const float4 vt = __ldg(Vd + static_cast<size_t>(kt) * (D / 4) + lane);
        const float4 wh = __ldg(U + static_cast<size_t>(jh) * (D / 4) + lane);
        const float4 wt = __ldg(U + static_cast<size_t>(jt) * (D / 4) + lane);
        const float4 xh = __ldg(Vd + static_cast<size_t>(lh) * (D / 4) + lane);
        const float4 xt = __ldg(Vd + static_cast<size_t>(lt) * (D / 4) + lane);
        Vec4<float>::store(out_h + static_cast<size_t>(p) * D + lane * 4,
                           make_float4(fast_tanh(uh.x + vh.x), fast_tanh(uh.y + vh.y), fast_tanh(uh.z + vh.z), fast_tanh(uh.w + vh.w)));
        Vec4<float>::store(out_t + static_cast<size_t>(p) * D + lane * 4,
                           make_float4(fast_tanh(ut.x + vt.x), fast_tanh(ut.y + vt.y), fast_tanh(ut.z + vt.z), fast_tanh(ut.w + vt.w)));
        if (two) {
            Vec4<float>::store(out_h + static_cast<size_t>(p2) * D + lane * 4,
                               make_float4(fast_tanh(wh.x + xh.x), fast_tanh(wh.y + xh.y), fast_tanh(wh.z + xh.z), fast_tanh(wh.w + xh.w)));
            Vec4<float>::store(out_t + static_cast<size_t>(p2) * D + lane * 4,
                               make_float4(fast_tanh(wt.x + xt.x), fast_tanh(wt.y + xt.y), fast_tanh(wt.z + xt.z), fast_tanh(wt.w + xt.w)));
        }
    }
}

// dpre = dout (1 - out^2) for both sides, float4 grid-stride
__global__ void __launch_bounds__(256)
tanh_bwd2_kernel(const float* __restrict__ dh, const float* __restrict__ dt, const float* __restrict__ oh,
                 const float* __restrict__ ot, float* __restrict__ ph, float* __restrict__ pt, long long count4) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < count4; i += stride) {
        const float4 a = Vec4<float>::load(dh + i * 4), b = Vec4<float>::load(oh + i * 4);
        Vec4<float>::store(ph + i * 4, make_float4(a.x * (1.f - b.x * b.x), a.y * (1.f - b.y * b.y),
                                                   a.z * (1.f - b.z * b.z), a.w * (1.f - b.w * b.w)));
        const float4 c = Vec4<float>::load(dt + i * 4), e = Vec4<float>::load(ot + i * 4);
        Vec4<float>::store(pt + i * 4, make_float4(c.x * (1.f - e.x * e.x), c.y * (1.f - e.y * e.y),
                                                   c.z * (1.f - e.z * e.z), c.w * (1.f - e.w * e.w)));
    }
}

int launch_pair_dense_fwd(const gcgcn_batch* bt, const float* U, const float* Vd, const int* h_idx, const int* t_idx,
                          const int* dis_h, const int* dis_t, float* out_h, float* out_t, cudaStream_t st) {
    if (bt->total_pairs == 0) return GCGCN_OK;
    const int blocks = static_cast<int>(std::min<long long>((bt->total_pairs + 7) / 8, static_cast<long long>(sm_count()) * 8));
    pair_dense_fwd_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(U), reinterpret_cast<const float4*>(Vd),
                                                  h_idx, t_idx, dis_h, dis_t, out_h, out_t, bt->total_pairs);
    GCGCN_CHECK_LAUNCH("pair_dense_fwd");
    return GCGCN_OK;
}

// dU[r] = sum of dpre over the pairs that read node row r, dVd[k] likewise for distance row k: the segmented sums of
// the gather backward (same kernels, 128 columns each) applied to dpre = dout (1 - out^2), which lives in `dpre`
// (caller-owned, 2 * total_pairs * 128 floats).
int launch_pair_dense_bwd(const gcgcn_batch* bt, const float* dout_h, const float* dout_t, const float* out_h,
                          const float* out_t, int dis_rows, const int* dis_h, const int* dis_t, float* dU, float* dVd,
                          float* dpre, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (bt->total_nodes == 0) return GCGCN_OK;
    float* ph = dpre;
    float* pt = dpre + static_cast<size_t>(bt->total_pairs) * D;
    const long long count4 = bt->total_pairs * (D / 4);
    if (count4 > 0) {
        const int blocks = static_cast<int>(std::min<long long>((count4 + 255) / 256, static_cast<long long>(sm_count()) * 16));
        tanh_bwd2_kernel<<<blocks, 256, 0, st>>>(dout_h, dout_t, out_h, out_t, ph, pt, count4);
        GCGCN_CHECK_LAUNCH("pair_dense_tanh_bwd");
    }
    pair_gather_bwd_feat_kernel<<<bt->total_nodes, 128, 0, st>>>(
        bt->node_ptr, reinterpret_cast<const long long*>(bt->pair_ptr), bt->row_doc, ph, pt, D / 4, D / 4, dU);
    GCGCN_CHECK_LAUNCH("pair_dense_bwd_rows");
    const int cells = dis_rows * D;
    long long warps = std::min<long long>(pair_dis_warps(), std::max<long long>(1, bt->total_pairs / 64));
    const int wpb = DIS_BWD_THREADS / WARP;
    const int blocks = ceil_div(warps, wpb);
    warps = static_cast<long long>(blocks) * wpb;
    const long long ppw = (bt->total_pairs + warps - 1) / warps;
    const size_t need = static_cast<size_t>(warps) * cells * sizeof(float);
    if (ws == nullptr || ws_bytes < need)
        return fail(GCGCN_ERR_WORKSPACE, "pair_dense_bwd: workspace %zu < %zu bytes", ws_bytes, need);
    const size_t smem = static_cast<size_t>(wpb) * cells * sizeof(float);
    if (smem > 48 * 1024)
        return fail(GCGCN_ERR_UNSUPPORTED, "pair_dense_bwd: distance table of %d rows too large", dis_rows);
    pair_gather_bwd_dis_kernel<<<blocks, DIS_BWD_THREADS, smem, st>>>(ph, pt, 0, D, dis_rows, dis_h, dis_t,
                                                                       bt->total_pairs, ppw, static_cast<float*>(ws));
    GCGCN_CHECK_LAUNCH("pair_dense_bwd_dis");
    return launch_reduce_partials(static_cast<const float*>(ws), static_cast<int>(warps), cells, dVd, cells, nullptr, st);
}

}  // namespace gcgcn
