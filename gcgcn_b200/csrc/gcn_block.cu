// CAGGC / MAGGC graph-block kernels for DocRED-sized graphs (n <= 64): one CTA per (document, head).
//
// This is the block form of the dense-connected GraphConv stack (G:63-80, G:97-120) with the attention
// step that feeds it fused in:
//   forward   P_h = softmax(q_h q_h^T / sqrt(d_h))            (MHA, G:133-142; key = query projection)
//             or the given attention map (CAGGC: GATAttention's map from the edge pass)
//             for l < L:  Z_l = Zx_l + g_{<l} Winner_l ;  g_l = relu((E_l + P Z_l) / r) ;  F_l = g_l + x_l
//   backward  dE_l = dN_l = relu'(g_l) dF_l / r ;  dA += dN_l Z_l^T ;  dZ_l = P^T dN_l ;  dense-connect push-down ;
//             then, without leaving the CTA, the softmax backward dS = P (dA - rowsum(dA P)) and either
//             dq_h = scale (dS + dS^T) q_h  (MHA)  or dS itself (GAT; the edge pass consumes it).
//
// Why a second generation of the stack kernel (gcn_stack_mma.cu): ncu showed the first one latency bound
// (long-scoreboard stalls on five dependent global round trips per CTA, 30-45 % of the SM's warps
// resident).  Here every global read a CTA needs is issued up front -- the per-sub-layer projection tiles
// with cp.async into a two-deep shared-memory ring, the epilogue operands into registers ahead of the
// MMA loop that precedes their use -- the attention map / its gradient never leave shared memory between
// sub-layers, and the softmax, its backward and the dq reduction ride along instead of being three more
// passes over [H][sum n^2] arrays in HBM.
//
// All small matrix products run on the tensor cores (mma.sync m16n8k8 TF32, 3xTF32 split, see
// mma_tf32.cuh).  Block configuration only: slab = in_dim = 128, ReLU + residual, no dropout masks
// (train-mode masks take the general path in gcn_stack_mma.cu).
#include "common.cuh"
#include "mma_tf32.cuh"

namespace gcgcn {

constexpr int BK_THREADS = 256;
constexpr int BK_WARPS = BK_THREADS / WARP;

enum { BK_OUT_DA = 0, BK_OUT_DS = 1, BK_OUT_DQ = 2 };

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Row softmax of the MHA scores of one (document, head): one warp per row, scores parked in the row of
// the shared attention tile they will occupy.  Writes P (global, rows < n), the zero-padded tile and the
// reciprocal row normalisers 1 / (rowsum + [rowsum == 0])  (G:47-49).
template <int DH>
__device__ __forceinline__ void mha_rows(const float* __restrict__ qs, float* __restrict__ As, float* __restrict__ rs,
                                         float* __restrict__ Pg, int n, int NP, int LDA, float scale) {
    constexpr int LQ = DH + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < NP; i += BK_WARPS) {
        float* row = As + i * LDA;
        if (i >= n) {
            for (int j = lane; j < NP; j += WARP) row[j] = 0.f;
            if (lane == 0) rs[i] = 1.f;
            continue;
        }
        float qi[DH];
#pragma unroll
        for (int k = 0; k < DH; ++k) qi[k] = qs[i * LQ + k];
        float m = -INFINITY;
        for (int j = lane; j < n; j += WARP) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < DH; ++k) s += qi[k] * qs[j * LQ + k];
            s *= scale;
            row[j] = s;
            m = fmaxf(m, s);
        }
        m = warp_max(m);
        float z = 0.f;
        for (int j = lane; j < n; j += WARP) {
            const float ex = expf(row[j] - m);
            row[j] = ex;
            z += ex;
        }
        z = warp_sum(z);
        float sum = 0.f;
        for (int j = lane; j < NP; j += WARP) {
            float p = 0.f;
            if (j < n) {
                p = row[j] / z;
                Pg[static_cast<size_t>(i) * n + j] = p;
            }
            row[j] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        if (lane == 0) rs[i] = 1.0f / (sum + (sum == 0.f ? 1.f : 0.f));
    }
}

// ---------------------------------------------------------------------------------------------
// DH = 0: attention map given in A.  DH > 0: MHA scores from the head slice of q (width DH); P is written.
template <int GD, int DH>
__global__ void __launch_bounds__(BK_THREADS, 3)
block_fwd_kernel(const int* __restrict__ node_ptr, const long long* __restrict__ pair_ptr,
                 const float* __restrict__ A, const float* __restrict__ q, float* __restrict__ P,
                 float* __restrict__ Z, const float* __restrict__ E, const float* __restrict__ Winner,
                 const float* __restrict__ x, float* __restrict__ G, float* __restrict__ F, int layers, int heads,
                 long long total_pairs, const int* __restrict__ doc_order, int first, float scale) {
    extern __shared__ __align__(16) float smem[];
    const int b = doc_order != nullptr ? doc_order[first + blockIdx.x] : blockIdx.x;
    const int h = blockIdx.y;
    const int node0 = node_ptr[b];
    const int n = node_ptr[b + 1] - node0;
    if (n == 0) return;
    constexpr int S = D, LDZ = GD + 8, CG = GD / 4, NGU = GD / 16;
    const int HD = heads * S, KI = (layers - 1) * GD;
    const int NP = (n + 15) & ~15, MT = NP / 16;
    const int LDA = NP + 4, LDG = KI + 4;

    float* As = smem;                  // [NP][LDA]   attention map, zero padded
    float* Zs0 = As + NP * LDA;        // [2][NP][LDZ] ring of projection tiles Zx_l (-> Z_l in place)
    float* Gs = Zs0 + 2 * NP * LDZ;    // [NP][LDG]   g_0 .. g_{L-2}
    float* rs = Gs + NP * LDG;         // [NP]        reciprocal row normalisers
    float* qs = Gs;                    // [n][DH+1]   head slice of q, dead before the first g_l is written

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const long long abase = static_cast<long long>(h) * total_pairs + pair_ptr[b];

    auto issue_z = [&](int l) {        // rows < n of the x-part projection of sub-layer l -> ring slot l & 1
        float* dst = Zs0 + (l & 1) * NP * LDZ;
        const float* src = Z + static_cast<size_t>(node0) * HD + static_cast<size_t>(h) * S + l * GD;
        for (int idx = tid; idx < n * CG; idx += BK_THREADS) {
            const int i = idx / CG, c4 = (idx - i * CG) * 4;
            cp_async16(dst + i * LDZ + c4, src + static_cast<size_t>(i) * HD + c4);
        }
    };
    issue_z(0);
    cp_async_commit();
    if (layers > 1) issue_z(1);
    cp_async_commit();
    // padding rows (never touched by cp.async or the epilogues) must be exact zeros: they are MMA operands
    for (int idx = tid; idx < (NP - n) * LDZ; idx += BK_THREADS) {
        Zs0[n * LDZ + idx] = 0.f;
        Zs0[NP * LDZ + n * LDZ + idx] = 0.f;
    }
    for (int idx = tid; idx < (NP - n) * LDG; idx += BK_THREADS) Gs[n * LDG + idx] = 0.f;

    if (DH > 0) {
        constexpr int LQ = DH + 1, Q4 = (DH > 0 ? DH : 4) / 4;
        for (int idx = tid; idx < n * Q4; idx += BK_THREADS) {
            const int j = idx / Q4, k4 = (idx - j * Q4) * 4;
            const float4 v = ld4g(q + static_cast<size_t>(node0 + j) * D + h * DH + k4);
            float* d = qs + j * LQ + k4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        __syncthreads();
        mha_rows<(DH > 0 ? DH : 4)>(qs, As, rs, P + abase, n, NP, LDA, scale);
    } else {
        const float* Ab = A + abase;
        for (int i = warp; i < NP; i += BK_WARPS) {
            float s = 0.f;
            for (int j = lane; j < NP; j += WARP) {
                const float v = (i < n && j < n) ? Ab[static_cast<size_t>(i) * n + j] : 0.f;
                As[i * LDA + j] = v;
                s += v;
            }
            s = warp_sum(s);
            if (lane == 0) rs[i] = 1.0f / (s + (s == 0.f ? 1.f : 0.f));
        }
    }

    for (int l = 0; l < layers; ++l) {
        float* Zs = Zs0 + (l & 1) * NP * LDZ;
        const int kin = l * GD;
        const size_t colbase = static_cast<size_t>(h) * S + l * GD;
        cp_async_wait<1>();            // this thread's copies of tile l have landed ...
        __syncthreads();               // ... and everyone's; also publishes As/rs (l = 0) and g_{l-1} (l > 0)
        if (l > 0) {
            // Z_l += g_{<l} Winner_l   (dense connection, row-local in the reference: G:72-73)
            const float* wsrc = Winner + (static_cast<size_t>(h) * layers + l) * S * GD;
            for (int u = warp; u < MT * NGU; u += BK_WARPS) {
                const int mt = u / NGU, ng = u - mt * NGU;
                float c[2][4];
                float* zc = Zs + (16 * mt) * LDZ + 16 * ng;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const float2 lo = *reinterpret_cast<const float2*>(zc + g * LDZ + 8 * nt + 2 * t);
                    const float2 hi = *reinterpret_cast<const float2*>(zc + (g + 8) * LDZ + 8 * nt + 2 * t);
                    c[nt][0] = lo.x; c[nt][1] = lo.y; c[nt][2] = hi.x; c[nt][3] = hi.y;
                }
                warp_gemm<2, false>(c, kin / 8, Gs + (16 * mt) * LDG, LDG, wsrc + 16 * ng, GD);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    *reinterpret_cast<float2*>(zc + g * LDZ + 8 * nt + 2 * t) = make_float2(c[nt][0], c[nt][1]);
                    *reinterpret_cast<float2*>(zc + (g + 8) * LDZ + 8 * nt + 2 * t) = make_float2(c[nt][2], c[nt][3]);
                }
            }
            __syncthreads();
            for (int idx = tid; idx < n * CG; idx += BK_THREADS) {       // final Z_l, saved for backward
                const int i = idx / CG, c4 = (idx - i * CG) * 4;
                *reinterpret_cast<float4*>(Z + static_cast<size_t>(node0 + i) * HD + colbase + c4) =
                    *reinterpret_cast<const float4*>(Zs + i * LDZ + c4);
            }
        }
        // out = (E + P Z_l) / r ; g_l = relu(out) ; F_l = g_l + x_l
        for (int u = warp; u < MT * NGU; u += BK_WARPS) {
            const int mt = u / NGU, ng = u - mt * NGU;
            // epilogue operands first: their latency hides behind the MMA loop
            float2 e2[2][2], x2[2][2];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int i = 16 * mt + g + 8 * half;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int col = 16 * ng + 8 * nt + 2 * t;
                    e2[half][nt] = make_float2(0.f, 0.f);
                    x2[half][nt] = make_float2(0.f, 0.f);
                    if (i < n) {
                        e2[half][nt] = ld2g(E + static_cast<size_t>(node0 + i) * HD + colbase + col);
                        x2[half][nt] = ld2g(x + static_cast<size_t>(node0 + i) * S + l * GD + col);
                    }
                }
            }
            float c[2][4];
            zero_frag<2>(c);
            warp_gemm<2, false>(c, NP / 8, As + (16 * mt) * LDA, LDA, Zs + 16 * ng, LDZ);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int i = 16 * mt + g + 8 * half;
                if (i >= n) continue;
                const float rinv = rs[i];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int col = 16 * ng + 8 * nt + 2 * t;
                    const size_t off = static_cast<size_t>(node0 + i) * HD + colbase + col;
                    float2 o;
                    o.x = fmaxf((e2[half][nt].x + c[nt][2 * half]) * rinv, 0.f);
                    o.y = fmaxf((e2[half][nt].y + c[nt][2 * half + 1]) * rinv, 0.f);
                    *reinterpret_cast<float2*>(G + off) = o;
                    if (l < layers - 1) *reinterpret_cast<float2*>(Gs + i * LDG + kin + col) = o;
                    *reinterpret_cast<float2*>(F + off) = make_float2(o.x + x2[half][nt].x, o.y + x2[half][nt].y);
                }
            }
        }
        __syncthreads();               // ring slot l & 1 is free again
        if (l + 2 < layers) issue_z(l + 2);
        cp_async_commit();             // (possibly empty) keeps "all but the newest group" == tile l+1
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// OUT = BK_OUT_DA: dA (head-major, like A) is written and the softmax backward is left to the caller.
// OUT = BK_OUT_DS: A is a softmax output; dS = A (dA - rowsum(dA A)) is written instead (GAT, one head).
// OUT = BK_OUT_DQ: as DS, then dq_h = scale (dS + dS^T) q_h is written to the head slice of dq [rows][128].
template <int GD, int DH, int OUT>
__global__ void __launch_bounds__(BK_THREADS, 3)
block_bwd_kernel(const int* __restrict__ node_ptr, const long long* __restrict__ pair_ptr,
                 const float* __restrict__ A, const float* __restrict__ q, const float* __restrict__ Z,
                 const float* __restrict__ G, const float* __restrict__ Winner, const float* __restrict__ dF,
                 float* __restrict__ dZ, float* __restrict__ dE, float* __restrict__ dOut, int layers, int heads,
                 long long total_pairs, const int* __restrict__ doc_order, int first, float scale) {
    extern __shared__ __align__(16) float smem[];
    const int b = doc_order != nullptr ? doc_order[first + blockIdx.x] : blockIdx.x;
    const int h = blockIdx.y;
    const int node0 = node_ptr[b];
    const int n = node_ptr[b + 1] - node0;
    if (n == 0) return;
    constexpr int S = D, LDN = GD + 12, CG = GD / 4, RPP = BK_THREADS / CG, NGU = GD / 16;
    const int HD = heads * S, KI = (layers - 1) * GD;
    const int NP = (n + 15) & ~15, MT = NP / 16, NT8 = NP / 8;
    const int LDA = NP + 4, LDG = KI + 4;

    float* Ats = smem;                 // [NP][LDA]  A transposed: Ats[j][i] = A[i][j]
    float* dAs = Ats + NP * LDA;       // [NP][LDA]  dA accumulated over the sub-layers
    float* dNs = dAs + NP * LDA;       // [NP][LDN]  dN_l = relu'(g_l) dG_l / r
    float* Ts = dNs + NP * LDN;        // [NP][LDN]  Z_l, then dZ_l
    float* dGs = Ts + NP * LDN;        // [NP][LDG]  dense-connect gradient parked for sub-layers < l
    float* rs = dGs + NP * LDG;        // [NP]
    float* drs = rs + NP;              // [NP]
    float* qs = dNs;                   // [n][DH+1]  (final phase only)
    float* rowbuf = Ts;                // [BK_WARPS][NP]  (final phase only)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const long long abase = static_cast<long long>(h) * total_pairs + pair_ptr[b];
    const float* Ab = A + abase;

    for (int i = warp; i < NP; i += BK_WARPS) {
        float s = 0.f;
        for (int j = lane; j < NP; j += WARP) {
            const float v = (i < n && j < n) ? Ab[static_cast<size_t>(i) * n + j] : 0.f;
            Ats[j * LDA + i] = v;
            s += v;
        }
        s = warp_sum(s);
        if (lane == 0) { rs[i] = 1.0f / (s + (s == 0.f ? 1.f : 0.f)); drs[i] = 0.f; }
    }
    __syncthreads();

    const int cg = tid % CG, rg = tid / CG, c0 = cg * 4;
    for (int l = layers - 1; l >= 0; --l) {
        const size_t colbase = static_cast<size_t>(h) * S + l * GD;
        const float* wsrc = Winner + (static_cast<size_t>(h) * layers + l) * S * GD;
        const bool first_layer = (l == layers - 1);
        // (a) row-local: dG_l -> dN_l (shared), dE_l (global), dr (shared); Z_l -> shared
        for (int i = rg; i < NP; i += RPP) {
            float4 dn = make_float4(0.f, 0.f, 0.f, 0.f), zl = dn;
            float drp = 0.f;
            if (i < n) {
                const size_t off = static_cast<size_t>(node0 + i) * HD + colbase + c0;
                float4 dg = ld4g(dF + off);
                const float4 g4 = ld4g(G + off);
                zl = ld4g(Z + off);
                if (!first_layer) {
                    const float4 s4 = *reinterpret_cast<const float4*>(dGs + i * LDG + l * GD + c0);
                    dg.x += s4.x; dg.y += s4.y; dg.z += s4.z; dg.w += s4.w;
                }
                dg.x = g4.x > 0.f ? dg.x : 0.f; dg.y = g4.y > 0.f ? dg.y : 0.f;
                dg.z = g4.z > 0.f ? dg.z : 0.f; dg.w = g4.w > 0.f ? dg.w : 0.f;
                const float rinv = rs[i];
                dn.x = dg.x * rinv; dn.y = dg.y * rinv; dn.z = dg.z * rinv; dn.w = dg.w * rinv;
                *reinterpret_cast<float4*>(dE + off) = dn;
                drp = -(dn.x * g4.x + dn.y * g4.y + dn.z * g4.z + dn.w * g4.w);
            }
            *reinterpret_cast<float4*>(dNs + i * LDN + c0) = dn;
            *reinterpret_cast<float4*>(Ts + i * LDN + c0) = zl;
#pragma unroll
            for (int o = CG / 2; o > 0; o >>= 1) drp += __shfl_xor_sync(0xffffffffu, drp, o);
            if (cg == 0 && i < n) drs[i] += drp;
        }
        __syncthreads();
        // (c) dA += dN_l Z_l^T
        for (int u = warp; u < MT * NT8; u += BK_WARPS) {
            const int mt = u / NT8, jt = u - mt * NT8;
            float c[1][4];
            zero_frag<1>(c);
            warp_gemm<1, true>(c, GD / 8, dNs + (16 * mt) * LDN, LDN, Ts + (8 * jt) * LDN, LDN);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float2* p = reinterpret_cast<float2*>(dAs + (16 * mt + g + 8 * half) * LDA + 8 * jt + 2 * t);
                float2 v = make_float2(c[0][2 * half], c[0][2 * half + 1]);
                if (!first_layer) { const float2 cur = *p; v.x += cur.x; v.y += cur.y; }
                *p = v;
            }
        }
        __syncthreads();
        // (b) dZ_l = A^T dN_l  -> Ts (for the dense-connect push-down) and global
        for (int u = warp; u < MT * NGU; u += BK_WARPS) {
            const int jt = u / NGU, ng = u - jt * NGU;
            float c[2][4];
            zero_frag<2>(c);
            warp_gemm<2, false>(c, NP / 8, Ats + (16 * jt) * LDA, LDA, dNs + 16 * ng, LDN);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int j = 16 * jt + g + 8 * half;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int col = 16 * ng + 8 * nt + 2 * t;
                    const float2 v = make_float2(c[nt][2 * half], c[nt][2 * half + 1]);
                    *reinterpret_cast<float2*>(Ts + j * LDN + col) = v;
                    if (j < n)
                        *reinterpret_cast<float2*>(dZ + static_cast<size_t>(node0 + j) * HD + colbase + col) = v;
                }
            }
        }
        __syncthreads();
        // push dZ_l through the dense connection: dG_m[j][c'] += sum_c dZ_l[j][c] * Wn_l[128 + m*GD + c'][c]
        for (int u = warp; u < l * MT * NGU; u += BK_WARPS) {
            const int m = u / (MT * NGU), rem = u - m * (MT * NGU);
            const int jt = rem / NGU, ng = rem - jt * NGU;
            float c[2][4];
            zero_frag<2>(c);
            warp_gemm<2, true>(c, GD / 8, Ts + (16 * jt) * LDN, LDN, wsrc + static_cast<size_t>(m * GD + 16 * ng) * GD, GD);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int j = 16 * jt + g + 8 * half;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    float2* p = reinterpret_cast<float2*>(dGs + j * LDG + m * GD + 16 * ng + 8 * nt + 2 * t);
                    float2 v = make_float2(c[nt][2 * half], c[nt][2 * half + 1]);
                    if (!first_layer) { const float2 cur = *p; v.x += cur.x; v.y += cur.y; }
                    *p = v;
                }
            }
        }
        __syncthreads();
    }

    // ---- attention gradient leaves the CTA -------------------------------------------------
    if (OUT == BK_OUT_DA) {
        float* dAb = dOut + abase;
        for (int i = warp; i < n; i += BK_WARPS) {
            const float dr = drs[i];
            for (int j = lane; j < n; j += WARP) dAb[static_cast<size_t>(i) * n + j] = dAs[i * LDA + j] + dr;
        }
        return;
    }
    if (OUT == BK_OUT_DQ) {
        constexpr int LQ = DH + 1, Q4 = (DH > 0 ? DH : 4) / 4;
        for (int idx = tid; idx < n * Q4; idx += BK_THREADS) {
            const int j = idx / Q4, k4 = (idx - j * Q4) * 4;
            const float4 v = ld4g(q + static_cast<size_t>(node0 + j) * D + h * DH + k4);
            float* d = qs + j * LQ + k4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    }
    // softmax backward, one warp per row: dS_ij = P_ij (dA_ij - sum_j dA_ij P_ij)
    for (int i = warp; i < n; i += BK_WARPS) {
        const float dr = drs[i];
        float dot = 0.f;
        for (int j = lane; j < n; j += WARP) dot += (dAs[i * LDA + j] + dr) * Ats[j * LDA + i];
        dot = warp_sum(dot);
        for (int j = lane; j < n; j += WARP) {
            const float ds = Ats[j * LDA + i] * (dAs[i * LDA + j] + dr - dot);
            if (OUT == BK_OUT_DS) dOut[abase + static_cast<size_t>(i) * n + j] = ds;
            else dAs[i * LDA + j] = ds;
        }
    }
    if (OUT == BK_OUT_DQ) {
        __syncthreads();
        // dq_h[i,:] = scale * sum_j (dS_ij + dS_ji) q_h[j,:]      (S = scale q q^T is symmetric in q)
        constexpr int DHH = DH > 0 ? DH : 4, LQ = DHH + 1, GROUPS = WARP / DHH;
        const int k = lane % DHH, grp = lane / DHH;
        float* buf = rowbuf + warp * NP;
        for (int i = warp; i < n; i += BK_WARPS) {
            for (int j = lane; j < n; j += WARP) buf[j] = dAs[i * LDA + j] + dAs[j * LDA + i];
            __syncwarp();
            float acc = 0.f;
            for (int j = grp; j < n; j += GROUPS) acc += buf[j] * qs[j * LQ + k];
            if (GROUPS >= 2) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            if (GROUPS >= 4) acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            if (grp == 0) dOut[static_cast<size_t>(node0 + i) * D + h * DHH + k] = acc * scale;
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
static size_t block_fwd_smem(int np, int layers, int gd) {
    const int ki = (layers - 1) * gd;
    return (static_cast<size_t>(np) * (np + 4) + 2 * static_cast<size_t>(np) * (gd + 8) +
            static_cast<size_t>(np) * (ki + 4) + np) * sizeof(float);
}
static size_t block_bwd_smem(int np, int layers, int gd) {
    const int ki = (layers - 1) * gd;
    return (2 * static_cast<size_t>(np) * (np + 4) + 2 * static_cast<size_t>(np) * (gd + 12) +
            static_cast<size_t>(np) * (ki + 4) + 2 * np) * sizeof(float);
}

// Block kernels cover: documents of <= 64 nodes, slab 128 cut into 2 x 64 or 4 x 32 sub-layers, and -- when
// the attention is computed in-kernel -- heads of width 16 or 32.
bool block_kernels_usable(const gcgcn_batch* bt, int heads, int layers, int slab, bool mha) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("GCGCN_STACK");
        enabled = (e != nullptr && e[0] != 0 && e[0] != 'b' && e[0] != 'B') ? 0 : 1;   // GCGCN_STACK=mma|simt disables
    }
    if (!enabled || slab != D || layers < 1 || slab % layers != 0) return false;
    const int gd = slab / layers;
    if (bt->max_nodes > 64 || !(gd == 32 || gd == 64)) return false;
    if (mha) {
        if (heads < 1 || D % heads != 0) return false;
        const int dh = D / heads;
        if (!(dh == 16 || dh == 32)) return false;
    }
    return true;
}

template <class LaunchFn>
static int for_each_size_class(const gcgcn_batch* bt, LaunchFn fn) {
    if (bt->doc_order == nullptr) return fn(bt->num_docs, 0, bt->max_nodes, static_cast<const int*>(nullptr));
    static const int cap[4] = {64, 48, 32, 16};
    int first = 0;
    for (int c = 0; c < 4; ++c) {
        const int count = bt->class_end[c] - first;
        if (count > 0) GCGCN_TRY(fn(count, first, cap[c] < bt->max_nodes ? cap[c] : bt->max_nodes, bt->doc_order));
        first = bt->class_end[c] > first ? bt->class_end[c] : first;
    }
    return GCGCN_OK;
}

template <typename K>
static int prepare_kernel(K kernel, size_t max_bytes, const char* name) {
    return cuda_ok(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(max_bytes)), name);
}

// A != nullptr: attention given (q, P unused).  A == nullptr: MHA from q, P written.
int launch_block_fwd(const gcgcn_batch* bt, int heads, int layers, const float* A, const float* q, float* P,
                     float* Z, const float* E, const float* Winner, const float* x, float* G, float* F,
                     cudaStream_t st) {
    if (bt->num_docs == 0) return GCGCN_OK;
    const int gd = D / layers;
    const int dh = A != nullptr ? 0 : D / heads;
    const float scale = dh > 0 ? 1.0f / sqrtf(static_cast<float>(dh)) : 1.f;
    const long long* pp = reinterpret_cast<const long long*>(bt->pair_ptr);
    const size_t max_smem = block_fwd_smem(64, layers, gd);
#define GCGCN_BK_FWD(GD_, DH_)                                                                                       \
    if (gd == GD_ && dh == DH_) {                                                                                    \
        static bool ready = false;                                                                                   \
        if (!ready) { GCGCN_TRY(prepare_kernel(block_fwd_kernel<GD_, DH_>, max_smem, "block_fwd")); ready = true; }  \
        return for_each_size_class(bt, [&](int count, int first, int nmax, const int* order) -> int {                \
            const size_t smem = block_fwd_smem((nmax + 15) & ~15, layers, gd);                                       \
            block_fwd_kernel<GD_, DH_><<<dim3(count, heads), BK_THREADS, smem, st>>>(                                \
                bt->node_ptr, pp, A, q, P, Z, E, Winner, x, G, F, layers, heads, bt->total_pairs, order, first,     \
                scale);                                                                                              \
            GCGCN_CHECK_LAUNCH(DH_ > 0 ? "block_fwd<mha>" : "block_fwd<given>");                                     \
            return GCGCN_OK;                                                                                         \
        });                                                                                                          \
    }
    GCGCN_BK_FWD(64, 0)
    GCGCN_BK_FWD(64, 16)
    GCGCN_BK_FWD(64, 32)
    GCGCN_BK_FWD(32, 0)
    GCGCN_BK_FWD(32, 16)
    GCGCN_BK_FWD(32, 32)
#undef GCGCN_BK_FWD
    return fail(GCGCN_ERR_UNSUPPORTED, "block_fwd: sub-layer width %d / head width %d not supported", gd, dh);
}

// out_mode: BK_OUT_DA -> dOut = dA [heads][total_pairs]; BK_OUT_DS -> dOut = dS (same shape; A must be a softmax
// output); BK_OUT_DQ -> dOut = dq [total_nodes][128] (A = the MHA probabilities, q their query projection).
int launch_block_bwd(const gcgcn_batch* bt, int heads, int layers, int out_mode, const float* A, const float* q,
                     const float* Z, const float* G, const float* Winner, const float* dF, float* dZ, float* dE,
                     float* dOut, cudaStream_t st) {
    if (bt->num_docs == 0) return GCGCN_OK;
    const int gd = D / layers;
    const int dh = out_mode == BK_OUT_DQ ? D / heads : 0;
    const float scale = dh > 0 ? 1.0f / sqrtf(static_cast<float>(dh)) : 1.f;
    const long long* pp = reinterpret_cast<const long long*>(bt->pair_ptr);
    const size_t max_smem = block_bwd_smem(64, layers, gd);
#define GCGCN_BK_BWD(GD_, DH_, OUT_, NAME_)                                                                          \
    if (gd == GD_ && dh == DH_ && out_mode == OUT_) {                                                                \
        static bool ready = false;                                                                                   \
        if (!ready) {                                                                                                \
            GCGCN_TRY(prepare_kernel(block_bwd_kernel<GD_, DH_, OUT_>, max_smem, "block_bwd"));                      \
            ready = true;                                                                                            \
        }                                                                                                            \
        return for_each_size_class(bt, [&](int count, int first, int nmax, const int* order) -> int {                \
            const size_t smem = block_bwd_smem((nmax + 15) & ~15, layers, gd);                                       \
            block_bwd_kernel<GD_, DH_, OUT_><<<dim3(count, heads), BK_THREADS, smem, st>>>(                          \
                bt->node_ptr, pp, A, q, Z, G, Winner, dF, dZ, dE, dOut, layers, heads, bt->total_pairs, order,      \
                first, scale);                                                                                       \
            GCGCN_CHECK_LAUNCH(NAME_);                                                                               \
            return GCGCN_OK;                                                                                         \
        });                                                                                                          \
    }
    GCGCN_BK_BWD(64, 0, BK_OUT_DA, "block_bwd<dA>")
    GCGCN_BK_BWD(32, 0, BK_OUT_DA, "block_bwd<dA>")
    GCGCN_BK_BWD(64, 0, BK_OUT_DS, "block_bwd<dS>")
    GCGCN_BK_BWD(32, 0, BK_OUT_DS, "block_bwd<dS>")
    GCGCN_BK_BWD(64, 16, BK_OUT_DQ, "block_bwd<dq>")
    GCGCN_BK_BWD(64, 32, BK_OUT_DQ, "block_bwd<dq>")
    GCGCN_BK_BWD(32, 16, BK_OUT_DQ, "block_bwd<dq>")
    GCGCN_BK_BWD(32, 32, BK_OUT_DQ, "block_bwd<dq>")
#undef GCGCN_BK_BWD
    return fail(GCGCN_ERR_UNSUPPORTED, "block_bwd: sub-layer width %d / head width %d / mode %d not supported", gd, dh,
                out_mode);
}

}  // namespace gcgcn
