// On-device expansion of the compact per-document wire format (gcgcn_b200/featurize.py) into the dense
// [n, n, S, L] pair-context tensors the reference builds on the host and copies to the GPU for every document
// (config/Config.py:180-205 -> 343-353: sen_matrix bool, pos_matrix_h / pos_matrix_t int64: 77 MB per document
// at n = 42, S = 5, L = 512, written by a triple Python loop).  The wire format ships one 9-int row per
// (edge, sentence slot); this kernel writes the same tensors at HBM speed.
//   row = (u, v, slot, s0, s1, h0, h1, t0, t1): sentence tokens [s0, s1), head mention [h0, h1), tail mention [t0, t1)
//   sen[u,v,slot,k] = 1,  pos_h[u,v,slot,k] = dis_plus + bucket(k; h0, h1),  pos_t likewise   for k in [s0, s1), k < L
//   bucket(k; a, b) = -dis2idx[a - k] if k < a;  dis2idx[k - b] if k > b;  0 otherwise        (C:187-201)
//   dis2idx[d] = 0, 1, 2, 2, 3, 3, 3, 3, 4, ... = floor(log2 d) + 1 capped at 10               (C:106-116)
// Everything outside the listed spans is zero (the caller-visible call zero-fills first).
#include "common.cuh"

namespace gcgcn {

__device__ __forceinline__ int dis_bucket(int d) {      // d >= 0
    return d == 0 ? 0 : min(10, 32 - __clz(d));
}
__device__ __forceinline__ long long signed_bucket(int k, int a, int b) {
    if (k < a) return -static_cast<long long>(dis_bucket(a - k));
    if (k > b) return dis_bucket(k - b);
    return 0;
}

__global__ void __launch_bounds__(128)
expand_pair_context_kernel(const int* __restrict__ slots, int num_slots, int n, int S, int L, int dis_plus,
                           unsigned char* __restrict__ sen, long long* __restrict__ pos_h,
                           long long* __restrict__ pos_t) {
    for (int e = blockIdx.x; e < num_slots; e += gridDim.x) {
        const int* r = slots + static_cast<size_t>(e) * 9;
        const int u = r[0], v = r[1], slot = r[2], s0 = r[3], s1 = min(r[4], L);
        if (slot >= S || u >= n || v >= n) continue;      // slots past max_num are truncated away (C:220-222)
        const size_t base = ((static_cast<size_t>(u) * n + v) * S + slot) * L;
        for (int k = s0 + threadIdx.x; k < s1; k += blockDim.x) {
            sen[base + k] = 1;
            pos_h[base + k] = dis_plus + signed_bucket(k, r[5], r[6]);
            pos_t[base + k] = dis_plus + signed_bucket(k, r[7], r[8]);
        }
    }
}

int launch_expand_pair_context(const int* slots, int num_slots, int n, int S, int L, int dis_plus, unsigned char* sen,
                               long long* pos_h, long long* pos_t, cudaStream_t st) {
    const size_t cells = static_cast<size_t>(n) * n * S * L;
    if (cells == 0) return GCGCN_OK;
    GCGCN_TRY(cuda_ok(cudaMemsetAsync(sen, 0, cells, st), "expand_pair_context: memset sen"));
    GCGCN_TRY(cuda_ok(cudaMemsetAsync(pos_h, 0, cells * sizeof(long long), st), "expand_pair_context: memset pos_h"));
    GCGCN_TRY(cuda_ok(cudaMemsetAsync(pos_t, 0, cells * sizeof(long long), st), "expand_pair_context: memset pos_t"));
    if (num_slots == 0) return GCGCN_OK;
    const int blocks = min(num_slots, sm_count() * 16);
    expand_pair_context_kernel<<<blocks, 128, 0, st>>>(slots, num_slots, n, S, L, dis_plus, sen, pos_h, pos_t);
    GCGCN_CHECK_LAUNCH("expand_pair_context");
    return GCGCN_OK;
}

}  // namespace gcgcn
