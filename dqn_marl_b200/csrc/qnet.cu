// Q-network act / learn — DQNNetwork.forward, DQNAgent.act and DQNAgent.learn of the reference
// (Louvre_Evacuation/agents/dqn_agent.py:15-61, :101-124, :126-168) as hand-written CUDA.
//
//   forward : conv1(6->32) conv2(32->64) conv3(64->128), 3x3 pad 1, ReLU; fc1(15488->512) ReLU Dropout(0.2);
//             fc2(512->256) ReLU; fc3(256->5).  The NHWC->NCHW permute of dqn_agent.py:37-45 never happens:
//             activations stay NHWC and the loaders of gemm_f32.cuh index them directly.
//   act     : fc3 + first-max argmax + epsilon-greedy keyed draw in one kernel (dqn_agent.py:103-104,124).
//   learn   : target forward -> max_a' Q^-(s',a');  online forward (activations kept);  fused TD target + loss +
//             dL/dq (dqn_agent.py:143-151);  hand-written backward;  global-norm clip folded into a fused
//             multi-tensor Adam (dqn_agent.py:158-160);  hard / Polyak target sync (:170-172).
//
// INTERNAL PARAMETER LAYOUTS (the 12 bound tensors; the Python side permutes at the state_dict boundary):
//   convK.weight  [(kh*3+kw)*Cin + c][Cout]      (PyTorch: [Cout][Cin][kh][kw])
//   fc1.weight    [512][p*128 + c], p = i*11+j   (PyTorch: [512][c*121 + p])
//   fc2.weight / fc3.weight / all biases: PyTorch layout.
// Gradients and Adam moments use the same layouts, so clip/Adam are layout-agnostic elementwise passes.
#include <cmath>
#include <cstdint>
#include <new>
#include "common.h"
#include "gemm_f32.cuh"
#include "gemm_tc.cuh"
#include "philox.cuh"
#include "qnet_bf16.cuh"

namespace mq {

constexpr int C1 = 32, C2 = 64, C3 = 128, CIN = 6, PIX = 121, FLAT = PIX * C3, H1 = 512, H2 = 256, NA = MQ_N_ACTIONS;

enum { P_C1W, P_C1B, P_C2W, P_C2B, P_C3W, P_C3B, P_F1W, P_F1B, P_F2W, P_F2B, P_F3W, P_F3B };
static const long long kParamCount[MQ_QNET_TENSORS] = {
    9LL * CIN * C1, C1, 9LL * C1 * C2, C2, 9LL * C2 * C3, C3, (long long)H1 * FLAT, H1, (long long)H2 * H1, H2, (long long)NA * H2, NA};

// ---- fc3 head ------------------------------------------------------------------------------------------
// one warp per sample: q[a] = h2 . W3[a] + b3[a]
// mode 0: write q.  mode 1 (act): epsilon-greedy -> action_out.  mode 2: q_sel[b] = q[action[b]].  mode 3: q_sel[b] = max_a q.
__global__ void __launch_bounds__(256)
qhead_kernel(const float* __restrict__ h2, const float* __restrict__ w3, const float* __restrict__ b3, long long B, int mode,
             float* __restrict__ q_out, const long long* __restrict__ action_in, float* __restrict__ q_sel,
             int* __restrict__ action_out, float eps, unsigned long long seed, unsigned env_id_base, unsigned tick, int n_robots) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float q[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) q[a] = 0.f;
    const float* h = h2 + b * H2;
    for (int k = lane; k < H2; k += 32) {
        const float hv = h[k];
#pragma unroll
        for (int a = 0; a < NA; ++a) q[a] = fmaf(hv, __ldg(w3 + a * H2 + k), q[a]);
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) {
#pragma unroll
        for (int o = 16; o; o >>= 1) q[a] += __shfl_xor_sync(0xFFFFFFFFu, q[a], o);
        q[a] += __ldg(b3 + a);
    }
    if (lane != 0) return;
    if (q_out) {
#pragma unroll
        for (int a = 0; a < NA; ++a) q_out[b * NA + a] = q[a];
    }
    int best = 0;
#pragma unroll
    for (int a = 1; a < NA; ++a) if (q[a] > q[best]) best = a;          // np.argmax: first maximum (dqn_agent.py:124)
    if (mode == 1) {
        // training and np.random.random() <= epsilon -> random.randrange(5) (dqn_agent.py:103-104), keyed draws
        const unsigned env = env_id_base + (unsigned)(b / n_robots), robot = (unsigned)(b % n_robots);
        const uint4 w = philox4x32(env, tick, robot, STREAM_AGENT, seed);
        const double u = u53(w.x, w.y);
        action_out[b] = (eps > 0.f && u <= (double)eps) ? (int)__umulhi(w.z, (unsigned)NA) : best;
    } else if (mode == 2) {
        q_sel[b] = q[(int)action_in[b]];
    } else if (mode == 3) {
        q_sel[b] = q[best];
    }
}

// ---- TD target + loss + dL/dq (dqn_agent.py:146-151), single CTA: deterministic reduction -------------------
__global__ void __launch_bounds__(1024)
td_loss_kernel(const float* __restrict__ q_sa, const float* __restrict__ maxq_next, const float* __restrict__ reward,
               const uint8_t* __restrict__ done, const long long* __restrict__ action, long long B, float gamma, int huber,
               float* __restrict__ dq, float* __restrict__ loss_out) {
    __shared__ float red[32];
    float local = 0.f;
    const float invB = 1.f / (float)B;
    for (long long b = threadIdx.x; b < B; b += blockDim.x) {
        // target = rewards + (gamma * next_q * ~dones)
        const float y = reward[b] + (gamma * maxq_next[b]) * (done[b] ? 0.f : 1.f);
        const float diff = q_sa[b] - y;
        float l, g;
        if (huber) { const float ad = fabsf(diff); l = ad <= 1.f ? 0.5f * diff * diff : ad - 0.5f; g = ad <= 1.f ? diff : copysignf(1.f, diff); }
        else { l = diff * diff; g = 2.f * diff; }                    // F.mse_loss, reduction='mean'
        local += l;
        const int a = (int)action[b];
#pragma unroll
        for (int k = 0; k < NA; ++k) dq[b * NA + k] = (k == a) ? g * invB : 0.f;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (threadIdx.x == 0) *loss_out = v * invB;
    }
}

// ---- fc3 backward: dW3[a][k] = sum_b dq[b][a] h2[b][k];  db3[a] = sum_b dq[b][a];  dh2 = (dq W3) * (h2 > 0) --------
constexpr int FC3_CHUNK = 64;      // batch rows per block of the fc3 weight gradient
__global__ void __launch_bounds__(256)
fc3_wgrad_kernel(const float* __restrict__ dq, const float* __restrict__ h2, long long B, float* __restrict__ part_w,
                 float* __restrict__ part_b) {
    // grid = (NA, chunks); thread k owns dW3[a][k] of this chunk; fixed b order inside a chunk: deterministic
    const int a = blockIdx.x, k = threadIdx.x;
    const long long b0 = (long long)blockIdx.y * FC3_CHUNK, b1 = b0 + FC3_CHUNK < B ? b0 + FC3_CHUNK : B;
    float acc = 0.f, accb = 0.f;
    for (long long b = b0; b < b1; ++b) {
        const float g = __ldg(dq + b * NA + a);
        if (g != 0.f) { acc = fmaf(g, __ldg(h2 + b * H2 + k), acc); accb += g; }
    }
    part_w[((size_t)blockIdx.y * NA + a) * H2 + k] = acc;
    if (k == 0) part_b[(size_t)blockIdx.y * NA + a] = accb;
}
__global__ void __launch_bounds__(256)
fc3_dgrad_kernel(const float* __restrict__ dq, const float* __restrict__ w3, const float* __restrict__ h2, long long B,
                 float* __restrict__ dh2, __nv_bfloat16* __restrict__ dh2b) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H2) return;
    const long long b = idx / H2;
    const int k = (int)(idx - b * H2);
    float v = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a) v = fmaf(__ldg(dq + b * NA + a), __ldg(w3 + a * H2 + k), v);
    v = h2[idx] > 0.f ? v : 0.f;
    dh2[idx] = v;
    if (dh2b) dh2b[idx] = __float2bfloat16(v);
}

// ---- bias gradients: out[n] = sum_m X[m][n], two deterministic stages ---------------------------------------
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ X, long long M, int N, int rows_per_block, float* __restrict__ partial) {
    const long long m0 = (long long)blockIdx.x * rows_per_block;
    const long long m1 = m0 + rows_per_block < M ? m0 + rows_per_block : M;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float acc = 0.f;
        for (long long m = m0; m < m1; ++m) acc += __ldg(X + m * N + n);
        partial[(size_t)blockIdx.x * N + n] = acc;
    }
}
// streaming form (N % VEC == 0, N / VEC <= 256): a thread owns one 16-byte column vector and every RG-th row of the block's
// row range, 4 rows in flight; the row groups are combined through shared memory in a fixed order (deterministic)
template <typename T> struct ColVec;
template <> struct ColVec<float> {
    static constexpr int VEC = 4;
    static __device__ __forceinline__ void add(float* a, const float* p) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
    }
};
template <> struct ColVec<bf::bf16> {
    static constexpr int VEC = 8;
    static __device__ __forceinline__ void add(float* a, const bf::bf16* p) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
            a[2 * k] += __low2float(b2); a[2 * k + 1] += __high2float(b2);
        }
    }
};
template <typename T>
__global__ void __launch_bounds__(256)
colsum_stream_kernel(const T* __restrict__ X, long long M, int N, long long rows_per_block, float* __restrict__ partial) {
    constexpr int VEC = ColVec<T>::VEC;
    __shared__ float red[256 * VEC];
    const int cols_v = N / VEC, RG = 256 / cols_v;
    const int cv = threadIdx.x % cols_v, rg = threadIdx.x / cols_v;
    const long long m0 = (long long)blockIdx.x * rows_per_block;
    const long long m1 = m0 + rows_per_block < M ? m0 + rows_per_block : M;
    float acc[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[u][k] = 0.f;
    if (rg < RG) {
        long long m = m0 + rg;
        for (; m + 3LL * RG < m1; m += 4LL * RG) {
#pragma unroll
            for (int u = 0; u < 4; ++u) ColVec<T>::add(acc[u], X + (m + (long long)u * RG) * N + cv * VEC);
        }
        for (; m < m1; m += RG) ColVec<T>::add(acc[0], X + m * N + cv * VEC);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) red[threadIdx.x * VEC + k] = (acc[0][k] + acc[1][k]) + (acc[2][k] + acc[3][k]);
    __syncthreads();
    if (threadIdx.x < cols_v) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float t = 0.f;
            for (int g = 0; g < RG; ++g) t += red[(g * cols_v + threadIdx.x) * VEC + k];
            partial[(size_t)blockIdx.x * N + threadIdx.x * VEC + k] = t;
        }
    }
}
// one warp per column: lanes stride over the partial rows, fixed shuffle tree: deterministic
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int blocks, int N, float* __restrict__ out) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    float acc = 0.f;
    for (int b = lane; b < blocks; b += 32) acc += partial[(size_t)b * N + n];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) out[n] = acc;
}

// ---- global grad norm (clip_grad_norm_, dqn_agent.py:158) + Adam (:85,160), multi-tensor ---------------------
struct TensorList {
    float* p[MQ_QNET_TENSORS]; float* g[MQ_QNET_TENSORS]; float* m[MQ_QNET_TENSORS]; float* v[MQ_QNET_TENSORS];
    float* t[MQ_QNET_TENSORS];
    long long n[MQ_QNET_TENSORS];
    long long chunk_start[MQ_QNET_TENSORS + 1];     // in units of CHUNK elements
};
constexpr int CHUNK = 4096;

__device__ __forceinline__ int find_tensor(const TensorList& tl, long long chunk) {
    int k = 0;
#pragma unroll
    for (int i = 1; i < MQ_QNET_TENSORS; ++i) if (chunk >= tl.chunk_start[i]) k = i;
    return k;
}

__global__ void __launch_bounds__(256)
sqnorm_partial_kernel(TensorList tl, float grad_scale, float* __restrict__ partial) {
    __shared__ float red[8];
    const long long chunk = blockIdx.x;
    const int k = find_tensor(tl, chunk);
    const long long base = (chunk - tl.chunk_start[k]) * CHUNK;
    const long long end = base + CHUNK < tl.n[k] ? base + CHUNK : tl.n[k];
    float acc = 0.f;
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x) { const float g = tl.g[k][i] * grad_scale; acc = fmaf(g, g, acc); }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { float s = 0.f; for (int w = 0; w < 8; ++w) s += red[w]; partial[chunk] = s; }
}
__global__ void __launch_bounds__(1024)
sqnorm_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ gnorm_out) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (threadIdx.x == 0) *gnorm_out = sqrtf(v);
    }
}
// PyTorch Adam (single-tensor formulation): m.lerp_(g, 1-b1); v = v*b2 + (1-b2) g*g;
// denom = sqrt(v)/sqrt(bc2) + eps;  p += -(lr/bc1) * m/denom.   g is first scaled by grad_scale (1/world after the
// all-reduce) and by clip_coef = min(1, clip/(norm + 1e-6)).
__global__ void __launch_bounds__(256)
clip_adam_kernel(TensorList tl, const float* __restrict__ gnorm, float grad_scale, float clip_norm, float beta1, float beta2,
                 float eps, float step_size, float bc2_sqrt) {
    const long long chunk = blockIdx.x;
    const int k = find_tensor(tl, chunk);
    const long long base = (chunk - tl.chunk_start[k]) * CHUNK;
    const long long end = base + CHUNK < tl.n[k] ? base + CHUNK : tl.n[k];
    float coef = clip_norm / (*gnorm + 1e-6f);
    coef = coef > 1.f ? 1.f : coef;
    const float gs = grad_scale * coef;
    float* __restrict__ P = tl.p[k]; float* __restrict__ M = tl.m[k]; float* __restrict__ V = tl.v[k];
    const float* __restrict__ G = tl.g[k];
    auto upd = [&](float g, float& p, float& m, float& v) {
        g *= gs;
        m = m + (g - m) * (1.f - beta1);
        v = v * beta2 + (1.f - beta2) * g * g;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p = p - step_size * (m / denom);
    };
    // 16-byte accesses on the aligned body of the chunk (every tensor of the flat buffer starts on a multiple of 4 floats)
    const bool vec = ((reinterpret_cast<uintptr_t>(P + base) | reinterpret_cast<uintptr_t>(M + base) | reinterpret_cast<uintptr_t>(V + base) |
                       reinterpret_cast<uintptr_t>(G + base)) & 15) == 0;
    const long long body = vec ? base + ((end - base) & ~3LL) : base;
    for (long long i = base + 4LL * threadIdx.x; i < body; i += 4LL * blockDim.x) {
        const float4 g4 = *reinterpret_cast<const float4*>(G + i);
        float4 p4 = *reinterpret_cast<float4*>(P + i), m4 = *reinterpret_cast<float4*>(M + i), v4 = *reinterpret_cast<float4*>(V + i);
        upd(g4.x, p4.x, m4.x, v4.x); upd(g4.y, p4.y, m4.y, v4.y); upd(g4.z, p4.z, m4.z, v4.z); upd(g4.w, p4.w, m4.w, v4.w);
        *reinterpret_cast<float4*>(P + i) = p4; *reinterpret_cast<float4*>(M + i) = m4; *reinterpret_cast<float4*>(V + i) = v4;
    }
    for (long long i = body + threadIdx.x; i < end; i += blockDim.x) {
        float p = P[i], m = M[i], v = V[i];
        upd(G[i], p, m, v);
        P[i] = p; M[i] = m; V[i] = v;
    }
}
__global__ void __launch_bounds__(256)
sync_target_kernel(TensorList tl, float tau) {
    const long long chunk = blockIdx.x;
    const int k = find_tensor(tl, chunk);
    const long long base = (chunk - tl.chunk_start[k]) * CHUNK;
    const long long end = base + CHUNK < tl.n[k] ? base + CHUNK : tl.n[k];
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x)
        tl.t[k][i] = tau >= 1.f ? tl.p[k][i] : tau * tl.p[k][i] + (1.f - tau) * tl.t[k][i];
}

// nn.Dropout(0.2) keep-mask (dqn_agent.py:33,57 — active in act() and learn() because the reference never
// calls .eval()): 16 keyed Bernoulli(1-p) bytes per Philox call.
__global__ void __launch_bounds__(256)
dropout_mask_kernel(uint8_t* __restrict__ mask, long long n, unsigned threshold, unsigned long long seed, unsigned long long counter) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // group of 4 bytes... 4 words -> 4 bytes
    const long long base = g * 4;
    if (base >= n) return;
    const uint4 w = philox4x32((unsigned)g, (unsigned)(g >> 32), (unsigned)counter, 64u + (unsigned)(counter >> 32), seed);
    const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) if (base + k < n) mask[base + k] = ws[k] >= threshold ? 1 : 0;   // keep with prob 1-p
}

}  // namespace mq

// =================================================================================================
struct mq_qnet {
    int device = 0, n_sms = 148;
    long long max_batch = 0;
    mq::TensorList tl;
    long long total_chunks = 0;
    // workspaces (device)
    float *a1 = nullptr, *a2 = nullptr, *a3 = nullptr, *h1 = nullptr, *h2 = nullptr;      // activations
    float *da1 = nullptr, *da2 = nullptr, *da3 = nullptr, *dh1 = nullptr, *dh2 = nullptr; // activation grads
    float *q = nullptr, *dq = nullptr, *q_sa = nullptr, *maxq = nullptr;
    float *partial = nullptr; size_t partial_cap = 0;
    float *norm_partial = nullptr, *gnorm = nullptr;
    int64_t launches = 0;
    // ---- bf16 tensor-core path (precision = 1) ----
    int precision = 0;                         // 0 = fp32 FFMA parity path, 1 = bf16 tcgen05 path
    // persistent convolutions that run with TWO sets of epilogue warps (bit 0 conv2 forward, 1 conv3 forward, 2 conv3 data
    // gradient, 3 conv2 data gradient); MQ_CONV_EPI8 overrides the default for A/B timing
    int conv_epi8 = 0;
    bool w_dirty[2] = {true, true};            // bf16 weight copies of [online, target] are stale
    mq::bf::bf16 *w2f[2] = {nullptr, nullptr}, *w3f[2] = {nullptr, nullptr}, *w1f[2] = {nullptr, nullptr};   // forward operands
    mq::bf::bf16 *w2d = nullptr, *w3d = nullptr, *w1t = nullptr;                                              // dgrad operands (online)
    mq::bf::bf16 *a1b = nullptr, *a2b = nullptr, *a3b = nullptr;                     // NHWC activations (TMA operands)
    mq::bf::bf16 *da3b = nullptr, *da2b = nullptr, *dh1b = nullptr, *dh2b = nullptr; // activation grads
    mq::bf::bf16 *h1b = nullptr, *wf2[2] = {nullptr, nullptr}, *wf2t = nullptr;      // fc2 operands: h1, W2 [256][512], W2^T [512][256]
    mq::bf::bf16 *A1 = nullptr, *w1c[2] = {nullptr, nullptr}, *da1b = nullptr;       // conv1: im2col rows [M][64], W [32][64], dY
};

namespace mq {

static void free_ws(mq_qnet* n) {
    float* ptrs[] = {n->a1, n->a2, n->a3, n->h1, n->h2, n->da1, n->da2, n->da3, n->dh1, n->dh2, n->q, n->dq, n->q_sa, n->maxq,
                     n->partial, n->norm_partial, n->gnorm};
    for (float* p : ptrs) cudaFree(p);
    bf::bf16* bptrs[] = {n->w2f[0], n->w2f[1], n->w3f[0], n->w3f[1], n->w1f[0], n->w1f[1], n->w2d, n->w3d, n->w1t, n->a1b, n->a2b,
                         n->a3b, n->da3b, n->da2b, n->dh1b, n->dh2b, n->h1b, n->wf2[0], n->wf2[1], n->wf2t, n->A1, n->w1c[0], n->w1c[1], n->da1b};
    for (bf::bf16* p : bptrs) cudaFree(p);
}

// ---- bf16 path ----------------------------------------------------------------------------------------------------
static int ew_blocks(long long total) { return (int)((total + 255) / 256); }

static cudaError_t alloc_bf16(mq_qnet* n) {
    if (n->a1b) return cudaSuccess;
    const size_t B = (size_t)n->max_batch, M = B * PIX, e = sizeof(bf::bf16);
    cudaError_t ce = cudaSuccess;
    auto alloc = [&](bf::bf16** p, size_t count) { if (ce == cudaSuccess) ce = cudaMalloc((void**)p, count * e); };
    for (int w = 0; w < 2; ++w) { alloc(&n->w2f[w], (size_t)C2 * 9 * C1); alloc(&n->w3f[w], (size_t)C3 * 9 * C2); alloc(&n->w1f[w], (size_t)H1 * FLAT); }
    alloc(&n->w2d, (size_t)C1 * 9 * C2); alloc(&n->w3d, (size_t)C2 * 9 * C3); alloc(&n->w1t, (size_t)FLAT * H1);
    alloc(&n->a1b, M * C1); alloc(&n->a2b, M * C2); alloc(&n->a3b, M * C3); alloc(&n->h1b, B * H1);
    for (int w = 0; w < 2; ++w) { alloc(&n->wf2[w], (size_t)H2 * H1); alloc(&n->w1c[w], (size_t)C1 * 64); }
    alloc(&n->A1, M * 64);
    if (n->tl.g[0]) { alloc(&n->da3b, M * C3); alloc(&n->da2b, M * C2); alloc(&n->dh1b, B * H1); alloc(&n->dh2b, B * H2); alloc(&n->wf2t, (size_t)H1 * H2); alloc(&n->da1b, M * C1); }
    return ce;
}

// fp32 master weights -> bf16 GEMM operands (after every optimizer step / target sync / load)
static void refresh_weights(mq_qnet* n, int which, cudaStream_t s) {
    if (!n->w_dirty[which]) return;
    float* const* W = which ? n->tl.t : n->tl.p;
    const bool bwd = which == 0 && n->tl.g[0];
    bf::conv_weight_bf16_kernel<<<ew_blocks(9 * C1 * C2), 256, 0, s>>>(W[P_C2W], n->w2f[which], bwd ? n->w2d : nullptr, C1, C2);
    bf::conv_weight_bf16_kernel<<<ew_blocks(9 * C2 * C3), 256, 0, s>>>(W[P_C3W], n->w3f[which], bwd ? n->w3d : nullptr, C2, C3);
    bf::cast_transpose_kernel<<<(H1 / 32) * ((FLAT + 31) / 32), 256, 0, s>>>(W[P_F1W], n->w1f[which], bwd ? n->w1t : nullptr, H1, FLAT);
    bf::cast_transpose_kernel<<<(H2 / 32) * ((H1 + 31) / 32), 256, 0, s>>>(W[P_F2W], n->wf2[which], bwd ? n->wf2t : nullptr, H2, H1);
    bf::conv1_weight_bf16_kernel<<<ew_blocks(C1 * 64), 256, 0, s>>>(W[P_C1W], n->w1c[which]);
    n->launches += 5;
    n->w_dirty[which] = false;
}

template <int BN>
static cudaError_t tc_gemm(mq_qnet* n, const bf::bf16* A, int lda, const bf::bf16* Bm, int ldb, int M, int N, int K, tc::Epilogue ep,
                           bool allow_split, cudaStream_t s) {
    int splits = 1;
    const long long tiles = (long long)((M + tc::BM - 1) / tc::BM) * ((N + BN - 1) / BN);
    if (allow_split && tiles < n->n_sms) {
        // fill one wave, never spill into a second: two CTAs per SM are co-resident (one with the 192 KB ring of BN = 256)
        splits = (int)(((BN == 256 ? 1LL : 2LL) * n->n_sms) / tiles);
        if (splits < 1) splits = 1;
        while (splits > 1 && (size_t)splits * M * N > n->partial_cap) --splits;
    }
    float* final_out = ep.out_f32;
    if (splits > 1) ep.partial = n->partial;
    cudaError_t e;
    if constexpr (BN == 256) e = tc::launch_pair<256, 6>(A, lda, Bm, ldb, M, N, K, ep, &splits, s);     // CTA pairs: 256 x 256 per cluster
    else e = tc::launch<BN, (BN == 128 ? 3 : 4)>(A, lda, Bm, ldb, M, N, K, ep, &splits, s);
    n->launches += 1;
    if (e == cudaSuccess && splits > 1) {
        GemmParams p{};
        p.M = M; p.N = N; p.C = final_out; p.Cb = ep.out_bf16; p.ldc = ep.ldc; p.partial = n->partial; p.splits = splits;
        p.bias = ep.bias; p.relu = ep.relu; p.drop = ep.drop; p.drop_scale = ep.drop_scale; p.mask_act = ep.mask_f32;
        size_t total = (size_t)M * N;
        int blocks = (int)((total + 255) / 256); if (blocks > 4 * n->n_sms) blocks = 4 * n->n_sms;
        splitk_epilogue_kernel<<<blocks, 256, 0, s>>>(p);
        n->launches += 1;
    }
    return e;
}

static void launch_colsum_bf16(mq_qnet* n, const bf::bf16* X, long long M, int N, float* out, cudaStream_t s);

// forward with every layer but the 5-output head on the tensor cores
// keep_im2col: the weight gradient of conv1 needs the im2col rows A1 afterwards (the online forward of a learn step); every
// other forward (target network, act, plain forward) builds the conv1 operand tile in shared memory and never writes A1
static cudaError_t forward_net_bf16(mq_qnet* n, int which, const float* obs, long long B, const uint8_t* drop_mask, cudaStream_t s,
                                    bool keep_im2col = false) {
    float* const* W = which ? n->tl.t : n->tl.p;
    const int M = (int)(B * PIX);
    refresh_weights(n, which, s);
    // conv1 (K = 54 padded to one 64-wide K-block; 495,616 x 32 x 64 at B = 4096 is an HBM stream, not a GEMM): persistent
    // kernel with the 4 KB weight tile resident; the operand rows are built in shared memory straight from the observation and,
    // when the backward pass will need them, also sent out to the im2col buffer A1 by a bulk tensor store of the same tile.
    tc::Epilogue ep{};
    ep.out_bf16 = n->a1b; ep.ldc = C1; ep.bias = W[P_C1B]; ep.relu = 1;
    cudaError_t e;
    e = tc::launch_conv1_obs<8>(obs, n->w1c[which], B, ep, n->n_sms, s, keep_im2col ? n->A1 : nullptr);
    n->launches += 1;
    if (e != cudaSuccess) return e;
    // conv2 / conv3: persistent implicit GEMMs (weights resident in shared memory), every tap a shifted zero-filled TMA box
    // of the NHWC activation (no im2col buffer)
    ep = tc::Epilogue{};
    ep.out_bf16 = n->a2b; ep.ldc = C2; ep.bias = W[P_C2B]; ep.relu = 1;
    e = (n->conv_epi8 & 1) ? tc::launch_conv_persistent<64, 32, 6, 8>(n->a1b, n->w2f[which], B, C1, C2, 0, ep, n->n_sms, s)
                           : tc::launch_conv_persistent<64, 32, 6>(n->a1b, n->w2f[which], B, C1, C2, 0, ep, n->n_sms, s);
    if (e != cudaSuccess) return e;
    ep = tc::Epilogue{};
    ep.out_bf16 = n->a3b; ep.ldc = C3; ep.bias = W[P_C3B]; ep.relu = 1;
    e = (n->conv_epi8 & 2) ? tc::launch_conv_persistent<128, 64, 3, 8>(n->a2b, n->w3f[which], B, C2, C3, 0, ep, n->n_sms, s)
                           : tc::launch_conv_persistent<128, 64, 3>(n->a2b, n->w3f[which], B, C2, C3, 0, ep, n->n_sms, s);
    if (e != cudaSuccess) return e;
    ep = tc::Epilogue{};
    ep.out_f32 = n->h1; ep.out_bf16 = n->h1b; ep.ldc = H1; ep.bias = W[P_F1B]; ep.relu = 1; ep.drop = drop_mask; ep.drop_scale = 1.f / (1.f - 0.2f);
    e = tc_gemm<256>(n, n->a3b, FLAT, n->w1f[which], FLAT, (int)B, H1, FLAT, ep, true, s);      // cta_group::2 pairs, 256 x 256: 1071 TFLOP/s (128 x 128: 881)
    if (e != cudaSuccess) return e;
    n->launches += 2;
    ep = tc::Epilogue{};
    ep.out_f32 = n->h2; ep.ldc = H2; ep.bias = W[P_F2B]; ep.relu = 1;
    e = tc_gemm<128>(n, n->h1b, H1, n->wf2[which], H1, (int)B, H2, H1, ep, false, s);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

// forward of one network over B samples; activations land in the handle's workspace
static int forward_net(mq_qnet* n, float* const* W, const float* obs, long long B, const uint8_t* drop_mask, cudaStream_t s) {
    const int M = (int)(B * PIX);
    GemmParams p{};
    p.batch = (int)B; p.partial = n->partial;
    // conv1: [M][54] x [54][32]
    p.M = M; p.N = C1; p.K = 9 * CIN; p.A = obs; p.B = W[P_C1W]; p.ldb = C1; p.C = n->a1; p.ldc = C1; p.bias = W[P_C1B]; p.relu = 1;
    n->launches += launch_gemm<A_IM2COL, B_ROW, 32, CIN>(p, n->partial_cap, n->n_sms, s);
    // conv2: [M][288] x [288][64]
    p.N = C2; p.K = 9 * C1; p.A = n->a1; p.B = W[P_C2W]; p.ldb = C2; p.C = n->a2; p.ldc = C2; p.bias = W[P_C2B];
    n->launches += launch_gemm<A_IM2COL, B_ROW, 64, C1>(p, n->partial_cap, n->n_sms, s);
    // conv3: [M][576] x [576][128]
    p.N = C3; p.K = 9 * C2; p.A = n->a2; p.B = W[P_C3W]; p.ldb = C3; p.C = n->a3; p.ldc = C3; p.bias = W[P_C3B];
    n->launches += launch_gemm<A_IM2COL, B_ROW, 128, C2>(p, n->partial_cap, n->n_sms, s);
    // fc1: [B][15488] x W1^T, ReLU, Dropout(0.2) (dqn_agent.py:56-57)
    p.M = (int)B; p.N = H1; p.K = FLAT; p.A = n->a3; p.lda = FLAT; p.B = W[P_F1W]; p.ldb = FLAT; p.C = n->h1; p.ldc = H1;
    p.bias = W[P_F1B]; p.drop = drop_mask; p.drop_scale = 1.f / (1.f - 0.2f);
    const bool small = B <= GEMM_SWAP_MAX_M;          // C1 (configs/dqn.yaml: B = 32) and single-env act(): transposed small-batch form
    if (small) n->launches += launch_gemm_swapped<A_ROW>(p, n->partial_cap, n->n_sms, s);
    else n->launches += launch_gemm<A_ROW, B_COL, 128, 1>(p, n->partial_cap, n->n_sms, s);
    // fc2
    p.N = H2; p.K = H1; p.A = n->h1; p.lda = H1; p.B = W[P_F2W]; p.ldb = H1; p.C = n->h2; p.ldc = H2; p.bias = W[P_F2B]; p.drop = nullptr;
    if (small) n->launches += launch_gemm_swapped<A_ROW>(p, n->partial_cap, n->n_sms, s);
    else n->launches += launch_gemm<A_ROW, B_COL, 64, 1>(p, n->partial_cap, n->n_sms, s);
    return 0;
}

template <typename T>
static void launch_colsum_t(mq_qnet* n, const T* X, long long M, int N, float* out, cudaStream_t s) {
    constexpr int VEC = ColVec<T>::VEC;
    if (N % VEC == 0 && N / VEC <= 256) {
        int blocks = 4 * n->n_sms;
        while (blocks > 1 && (size_t)blocks * N > n->partial_cap) blocks /= 2;
        long long rows = (M + blocks - 1) / blocks;
        if (rows < 64) rows = 64;
        blocks = (int)((M + rows - 1) / rows);
        colsum_stream_kernel<T><<<blocks, 256, 0, s>>>(X, M, N, rows, n->partial);
        colsum_final_kernel<<<(N * 32 + 255) / 256, 256, 0, s>>>(n->partial, blocks, N, out);
        n->launches += 2;
        return;
    }
    int rows = 256;
    int blocks = (int)((M + rows - 1) / rows);
    while ((size_t)blocks * N > n->partial_cap) { rows *= 2; blocks = (int)((M + rows - 1) / rows); }
    if (sizeof(T) == 4) colsum_partial_kernel<<<blocks, 256, 0, s>>>((const float*)X, M, N, rows, n->partial);
    else bf::colsum_partial_bf16_kernel<<<blocks, 256, 0, s>>>((const bf::bf16*)X, M, N, rows, n->partial);
    colsum_final_kernel<<<(N * 32 + 255) / 256, 256, 0, s>>>(n->partial, blocks, N, out);
    n->launches += 2;
}
static void launch_colsum(mq_qnet* n, const float* X, long long M, int N, float* out, cudaStream_t s) { launch_colsum_t<float>(n, X, M, N, out, s); }
static void launch_colsum_bf16(mq_qnet* n, const bf::bf16* X, long long M, int N, float* out, cudaStream_t s) { launch_colsum_t<bf::bf16>(n, X, M, N, out, s); }

// conv weight gradient on the tensor cores: split over samples so that ~one wave of CTAs runs, deterministic reduce
template <int BN, int STAGES, int AW>
static cudaError_t conv_wgrad(mq_qnet* n, const bf::bf16* X, const bf::bf16* dY, long long B, int Cin, int Cout, float* out, float* bias_out,
                              cudaStream_t s) {
    const int M = 9 * Cin;
    const int tiles = ((M + tc::BM - 1) / tc::BM) * ((Cout + BN - 1) / BN);
    int splits = n->n_sms / tiles;            // one CTA per SM (192 KB of shared memory each): never more CTAs than SMs
    if (splits < 1) splits = 1;
    // the last COLSUM_RESERVE floats of the partial buffer hold the per-split column sums of dY (the bias gradient), which
    // the kernel produces through a spare A row of ones when the shape has one
    constexpr size_t COLSUM_RESERVE = (size_t)1024 * 256;
    while (splits > 1 && ((size_t)splits * M * Cout > n->partial_cap - COLSUM_RESERVE || (size_t)splits * Cout > COLSUM_RESERVE)) --splits;
    tc::Epilogue ep{};
    ep.out_f32 = out; ep.ldc = Cout; ep.partial = n->partial;
    ep.colsum_partial = n->partial + (n->partial_cap - COLSUM_RESERVE);
    bool fused = false;
    cudaError_t e = tc::launch_conv_wgrad<BN, STAGES, AW>(X, dY, B, Cin, Cout, ep, &splits, s, &fused);
    n->launches += 1;
    if (e != cudaSuccess) return e;
    if (fused) {
        colsum_final_kernel<<<(Cout * 32 + 255) / 256, 256, 0, s>>>(ep.colsum_partial, splits, Cout, bias_out);
        n->launches += 1;
    }
    if (splits > 1) {
        GemmParams p{};
        p.M = M; p.N = Cout; p.C = out; p.ldc = Cout; p.partial = n->partial; p.splits = splits;
        size_t total = (size_t)M * Cout;
        int blocks = (int)((total + 255) / 256); if (blocks > 4 * n->n_sms) blocks = 4 * n->n_sms;
        splitk_epilogue_kernel<<<blocks, 256, 0, s>>>(p);
        n->launches += 1;
    }
    if (!fused) launch_colsum_bf16(n, dY, B * PIX, Cout, bias_out, s);      // after the split reduce: it reuses the partial buffer
    return e;
}

// backward of the bf16 path.  Needs: forward_net_bf16(online) just ran (A1, a1b, a2b, a3b, h1, h2 hold the online
// activations) and n->dq holds dL/dq.
// part: 0 = everything; 1 = head only (fc3, fc2, fc1: the last ~99 % of the flat gradient, so that its all-reduce can
// overlap the convolution backward); 2 = the convolution layers only (after a part-1 call)
static cudaError_t backward_bf16(mq_qnet* n, const float* state, long long B, const uint8_t* drop_online, cudaStream_t s, int part = 0) {
    float* const* W = n->tl.p; float* const* G = n->tl.g;
    const int M = (int)(B * PIX);
    cudaError_t e;
    tc::Epilogue ep{};
    GemmParams p{};
    if (part != 2) {
    {
        const int chunks = (int)((B + FC3_CHUNK - 1) / FC3_CHUNK);
        float* pw = n->partial; float* pb = n->partial + (size_t)chunks * NA * H2;
        fc3_wgrad_kernel<<<dim3(NA, chunks), H2, 0, s>>>(n->dq, n->h2, B, pw, pb);
        colsum_final_kernel<<<(NA * H2 * 32 + 255) / 256, 256, 0, s>>>(pw, chunks, NA * H2, G[P_F3W]);
        colsum_final_kernel<<<(NA * 32 + 255) / 256, 256, 0, s>>>(pb, chunks, NA, G[P_F3B]);
        n->launches += 2;
    }
    fc3_dgrad_kernel<<<(int)((B * H2 + 255) / 256), 256, 0, s>>>(n->dq, W[P_F3W], n->h2, B, n->dh2, n->dh2b);
    n->launches += 2;
    // fc2 on the tensor cores: dW2[256][512] = dh2^T h1 (MN-major operands, split over the batch) ; db2 ; dh1 = dh2 W2 masked
    // by relu(fc1) > 0 and the dropout mask
    {
        int splits = 16;
        while (splits > 1 && (size_t)splits * H2 * H1 > n->partial_cap) --splits;
        ep.out_f32 = G[P_F2W]; ep.ldc = H1; ep.partial = n->partial;
        if ((e = tc::launch_tn<128, 3>(n->dh2b, H2, n->h1b, H1, H2, H1, (int)B, ep, &splits, s)) != cudaSuccess) return e;
        n->launches += 1;
        if (splits > 1) {
            p = GemmParams{};
            p.M = H2; p.N = H1; p.C = G[P_F2W]; p.ldc = H1; p.partial = n->partial; p.splits = splits;
            splitk_epilogue_kernel<<<(H2 * H1 + 255) / 256, 256, 0, s>>>(p);
            n->launches += 1;
        }
    }
    launch_colsum(n, n->dh2, B, H2, G[P_F2B], s);
    ep = tc::Epilogue{};
    ep.out_f32 = n->dh1; ep.out_bf16 = n->dh1b; ep.ldc = H1; ep.mask_f32 = n->h1; ep.drop = drop_online; ep.drop_scale = 1.f / (1.f - 0.2f);
    if ((e = tc_gemm<128>(n, n->dh2b, H2, n->wf2t, H2, (int)B, H1, H2, ep, false, s)) != cudaSuccess) return e;
    launch_colsum(n, n->dh1, B, H1, G[P_F1B], s);
    // fc1 on the tensor cores: dW1 = dh1^T a3 (both operands MN-major: no transposes) ; da3 = dh1 W1 (masked by a3 > 0)
    ep = tc::Epilogue{};
    ep.out_f32 = G[P_F1W]; ep.ldc = FLAT;
    {
        int one = 1;
        if ((e = tc::launch_tn<128, 3>(n->dh1b, H1, n->a3b, FLAT, H1, FLAT, (int)B, ep, &one, s)) != cudaSuccess) return e;
        n->launches += 1;
    }
    ep = tc::Epilogue{};
    ep.out_bf16 = n->da3b; ep.ldc = FLAT; ep.mask_bf16 = n->a3b;
    if ((e = tc_gemm<128>(n, n->dh1b, H1, n->w1t, H1, (int)B, FLAT, H1, ep, false, s)) != cudaSuccess) return e;
    }
    if (part == 1) return cudaGetLastError();
    // conv3: dWc3[(t,c)][n] = im2col(a2)^T dY (implicit, split over samples) ; db ; da2 = conv_flip(dY, Wd3) (masked by a2 > 0).
    if ((e = conv_wgrad<128, 3, 64>(n, n->a2b, n->da3b, B, C2, C3, G[P_C3W], G[P_C3B], s)) != cudaSuccess) return e;
    ep = tc::Epilogue{};
    ep.out_bf16 = n->da2b; ep.ldc = C2; ep.mask_bf16 = n->a2b;
    e = (n->conv_epi8 & 4) ? tc::launch_conv_persistent<64, 64, 3, 8>(n->da3b, n->w3d, B, C3, C2, 1, ep, n->n_sms, s)
                           : tc::launch_conv_persistent<64, 64, 3>(n->da3b, n->w3d, B, C3, C2, 1, ep, n->n_sms, s);
    if (e != cudaSuccess) return e;
    // conv2
    if ((e = conv_wgrad<64, 4, 32>(n, n->a1b, n->da2b, B, C1, C2, G[P_C2W], G[P_C2B], s)) != cudaSuccess) return e;
    ep = tc::Epilogue{};
    ep.out_bf16 = n->da1b; ep.ldc = C1; ep.mask_bf16 = n->a1b;
    e = (n->conv_epi8 & 8) ? tc::launch_conv_persistent<32, 64, 4, 8>(n->da2b, n->w2d, B, C2, C1, 1, ep, n->n_sms, s)
                           : tc::launch_conv_persistent<32, 64, 4>(n->da2b, n->w2d, B, C2, C1, 1, ep, n->n_sms, s);
    if (e != cudaSuccess) return e;
    n->launches += 2;
    // conv1: dWc1^T [32][64] = dY^T A1 (MN-major operands, 32-wide A slabs, split over the rows), reduced and transposed into
    // the [(tap, c)][32] layout; column 54 of A1 is all ones, so row 54 of the product is db = the column sums of dY.
    // The observation needs no gradient.
    {
        int splits = n->n_sms;
        while (splits > 1 && (size_t)splits * C1 * 64 > n->partial_cap) --splits;
        ep = tc::Epilogue{};
        ep.out_f32 = n->partial; ep.ldc = 64; ep.partial = n->partial;
        if ((e = tc::launch_tn<64, 4, 32>(n->da1b, C1, n->A1, 64, C1, 64, M, ep, &splits, s)) != cudaSuccess) return e;
        if (splits == 1) return cudaErrorInvalidValue;      // (never: M / 64 k-blocks >> 1) the reduce below expects partials
        bf::conv1_wgrad_reduce_kernel<<<55, 256, 0, s>>>(n->partial, splits, G[P_C1W], G[P_C1B]);
        n->launches += 2;
    }
    return cudaGetLastError();
}

}  // namespace mq

extern "C" int mq_qnet_create(mq_qnet** out, int32_t device, int64_t max_batch, const mq_qnet_bind* bind) {
    MQ_REQUIRE(out && bind && max_batch > 0, "mq_qnet_create: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return mq::fail(MQ_ERR_CUDA, "mq_qnet_create: no CUDA device (this build has no CPU fallback)");
    MQ_REQUIRE(device >= 0 && device < ndev, "mq_qnet_create: device %d outside 0..%d", device, ndev - 1);
    MQ_ON_DEVICE(device);
    mq_qnet* n = new (std::nothrow) mq_qnet();
    if (!n) return mq::fail(MQ_ERR_ALLOC, "mq_qnet_create: out of host memory");
    n->device = device; n->max_batch = max_batch;
    if (const char* v = getenv("MQ_CONV_EPI8")) n->conv_epi8 = atoi(v);
    cudaDeviceGetAttribute(&n->n_sms, cudaDevAttrMultiProcessorCount, device);
    long long chunk = 0;
    for (int k = 0; k < MQ_QNET_TENSORS; ++k) {
        MQ_REQUIRE(bind->online[k] && bind->target[k], "mq_qnet_create: parameter tensor %d missing", k);
        n->tl.p[k] = bind->online[k]; n->tl.t[k] = bind->target[k]; n->tl.g[k] = bind->grad[k];
        n->tl.m[k] = bind->adam_m[k]; n->tl.v[k] = bind->adam_v[k];
        n->tl.n[k] = mq::kParamCount[k];
        n->tl.chunk_start[k] = chunk;
        chunk += (mq::kParamCount[k] + mq::CHUNK - 1) / mq::CHUNK;
    }
    n->tl.chunk_start[MQ_QNET_TENSORS] = chunk;
    n->total_chunks = chunk;
    const size_t B = (size_t)max_batch, f = sizeof(float);
    n->partial_cap = (size_t)16 << 20;     // 16 Mi floats = 64 MB of split-K / column-sum partials
    cudaError_t ce = cudaSuccess;
    auto alloc = [&](float** p, size_t count) { if (ce == cudaSuccess) ce = cudaMalloc((void**)p, count * f); };
    alloc(&n->a1, B * mq::PIX * mq::C1); alloc(&n->a2, B * mq::PIX * mq::C2); alloc(&n->a3, B * mq::FLAT);
    alloc(&n->h1, B * mq::H1); alloc(&n->h2, B * mq::H2);
    if (bind->grad[0]) {
        alloc(&n->da1, B * mq::PIX * mq::C1); alloc(&n->da2, B * mq::PIX * mq::C2); alloc(&n->da3, B * mq::FLAT);
        alloc(&n->dh1, B * mq::H1); alloc(&n->dh2, B * mq::H2);
    }
    alloc(&n->q, B * mq::NA); alloc(&n->dq, B * mq::NA); alloc(&n->q_sa, B); alloc(&n->maxq, B);
    alloc(&n->partial, n->partial_cap); alloc(&n->norm_partial, (size_t)chunk + 1); alloc(&n->gnorm, 1);
    if (ce != cudaSuccess) {
        mq::free_ws(n); delete n;
        return mq::fail(MQ_ERR_ALLOC, "mq_qnet_create: workspace allocation for max_batch=%lld failed: %s", (long long)max_batch,
                        cudaGetErrorString(ce));
    }
    *out = n;
    return MQ_OK;
}

extern "C" int mq_qnet_destroy(mq_qnet* n) {
    if (!n) return MQ_OK;
    MQ_ON_DEVICE(n->device);
    mq::free_ws(n);
    delete n;
    return MQ_OK;
}
extern "C" int64_t mq_qnet_launch_count(const mq_qnet* n) { return n ? n->launches : 0; }

extern "C" int mq_qnet_forward(mq_qnet* n, int32_t which, const float* obs, int64_t B, const uint8_t* drop_mask, float* q_out,
                               void* stream) {
    MQ_REQUIRE(n && obs && q_out, "mq_qnet_forward: null argument");
    MQ_REQUIRE(B > 0 && B <= n->max_batch, "mq_qnet_forward: batch %lld outside 1..%lld", (long long)B, n->max_batch);
    MQ_ON_DEVICE(n->device);
    cudaStream_t s = (cudaStream_t)stream;
    float* const* W = which ? n->tl.t : n->tl.p;
    if (n->precision == 1) {
        cudaError_t e = mq::forward_net_bf16(n, which ? 1 : 0, obs, B, drop_mask, s);
        if (e != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_qnet_forward (bf16 path): %s", cudaGetErrorString(e));
    } else {
        mq::forward_net(n, W, obs, B, drop_mask, s);
    }
    const int blocks = (int)((B * 32 + 255) / 256);
    mq::qhead_kernel<<<blocks, 256, 0, s>>>(n->h2, W[mq::P_F3W], W[mq::P_F3B], B, 0, q_out, nullptr, nullptr, nullptr, 0.f, 0, 0, 0, 1);
    n->launches += 1;
    MQ_CUDA(cudaGetLastError());
    return MQ_OK;
}

extern "C" int mq_qnet_act(mq_qnet* n, const float* obs, int64_t B, float eps, uint64_t seed, uint32_t env_id_base, uint32_t tick,
                           int32_t n_robots, const uint8_t* drop_mask, int32_t* action_out, float* q_out, void* stream) {
    MQ_REQUIRE(n && obs && action_out && n_robots >= 1, "mq_qnet_act: bad argument");
    MQ_REQUIRE(B > 0 && B <= n->max_batch, "mq_qnet_act: batch %lld outside 1..%lld", (long long)B, n->max_batch);
    MQ_ON_DEVICE(n->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (n->precision == 1) {
        cudaError_t e = mq::forward_net_bf16(n, 0, obs, B, drop_mask, s);
        if (e != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_qnet_act (bf16 path): %s", cudaGetErrorString(e));
    } else {
        mq::forward_net(n, n->tl.p, obs, B, drop_mask, s);
    }
    const int blocks = (int)((B * 32 + 255) / 256);
    mq::qhead_kernel<<<blocks, 256, 0, s>>>(n->h2, n->tl.p[mq::P_F3W], n->tl.p[mq::P_F3B], B, 1, q_out, nullptr, nullptr, action_out, eps,
                                            seed, env_id_base, tick, n_robots);
    n->launches += 1;
    MQ_CUDA(cudaGetLastError());
    return MQ_OK;
}

extern "C" int mq_qnet_explore_draw(float eps, uint64_t seed, uint32_t env, uint32_t tick, uint32_t robot, int32_t* action_out) {
    // the same words and comparison as qhead_kernel's mode 1, evaluated on the host
    const uint4 w = mq::philox4x32(env, tick, robot, mq::STREAM_AGENT, seed);
    const double u = mq::u53(w.x, w.y);
    if (eps > 0.f && u <= (double)eps) {
        if (action_out) *action_out = (int32_t)(((uint64_t)w.z * (uint64_t)mq::NA) >> 32);
        return 1;
    }
    return 0;
}

// backward of the fp32 parity path from n->dq (loss.backward(), dqn_agent.py:154-155); the online activations are in the workspace
static void backward_fp32(mq_qnet* n, const float* state, long long B, const uint8_t* drop_online, cudaStream_t s, int part = 0) {
    using namespace mq;
    float* const* W = n->tl.p; float* const* G = n->tl.g;
    GemmParams p{};
    if (part != 2) {
    {
        const int chunks = (int)((B + FC3_CHUNK - 1) / FC3_CHUNK);
        float* pw = n->partial; float* pb = n->partial + (size_t)chunks * NA * H2;
        fc3_wgrad_kernel<<<dim3(NA, chunks), H2, 0, s>>>(n->dq, n->h2, B, pw, pb);
        colsum_final_kernel<<<(NA * H2 * 32 + 255) / 256, 256, 0, s>>>(pw, chunks, NA * H2, G[P_F3W]);
        colsum_final_kernel<<<(NA * 32 + 255) / 256, 256, 0, s>>>(pb, chunks, NA, G[P_F3B]);
        n->launches += 2;
    }
    fc3_dgrad_kernel<<<(int)((B * H2 + 255) / 256), 256, 0, s>>>(n->dq, W[P_F3W], n->h2, B, n->dh2, nullptr);
    n->launches += 2;
    p.batch = (int)B; p.partial = n->partial;
    // fc2: dW2[256][512] = dh2^T h1 ; db2 ; dh1 = dh2 W2, masked by relu(fc1) > 0 and the dropout mask
    p.M = H2; p.N = H1; p.K = (int)B; p.A = n->dh2; p.lda = H2; p.B = n->h1; p.ldb = H1; p.C = G[P_F2W]; p.ldc = H1;
    n->launches += launch_gemm<A_COL, B_ROW, 128, 1>(p, n->partial_cap, n->n_sms, s);
    launch_colsum(n, n->dh2, B, H2, G[P_F2B], s);
    p = GemmParams{}; p.batch = (int)B; p.partial = n->partial;
    p.M = (int)B; p.N = H1; p.K = H2; p.A = n->dh2; p.lda = H2; p.B = W[P_F2W]; p.ldb = H1; p.C = n->dh1; p.ldc = H1;
    p.mask_act = n->h1; p.drop = drop_online; p.drop_scale = 1.f / (1.f - 0.2f);
    const bool small = B <= GEMM_SWAP_MAX_M;          // as in forward_net
    if (small) n->launches += launch_gemm_swapped<A_COL>(p, n->partial_cap, n->n_sms, s);
    else n->launches += launch_gemm<A_ROW, B_ROW, 128, 1>(p, n->partial_cap, n->n_sms, s);
    // fc1: dW1[512][15488] = dh1^T a3 ; db1 ; da3 = dh1 W1 masked by a3 > 0
    p = GemmParams{}; p.batch = (int)B; p.partial = n->partial;
    p.M = H1; p.N = FLAT; p.K = (int)B; p.A = n->dh1; p.lda = H1; p.B = n->a3; p.ldb = FLAT; p.C = G[P_F1W]; p.ldc = FLAT;
    n->launches += launch_gemm<A_COL, B_ROW, 128, 1>(p, n->partial_cap, n->n_sms, s);
    launch_colsum(n, n->dh1, B, H1, G[P_F1B], s);
    p = GemmParams{}; p.batch = (int)B; p.partial = n->partial;
    p.M = (int)B; p.N = FLAT; p.K = H1; p.A = n->dh1; p.lda = H1; p.B = W[P_F1W]; p.ldb = FLAT; p.C = n->da3; p.ldc = FLAT; p.mask_act = n->a3;
    if (small) n->launches += launch_gemm_swapped<A_COL>(p, n->partial_cap, n->n_sms, s);
    else n->launches += launch_gemm<A_ROW, B_ROW, 128, 1>(p, n->partial_cap, n->n_sms, s);
    }
    if (part == 1) return;
    // conv3: dWc3[(tap,c)][n] = im2col(a2)^T da3 ; db ; da2 = dgrad masked by a2 > 0
    const int M = (int)(B * PIX);
    p = GemmParams{}; p.batch = (int)B; p.partial = n->partial;
    p.M = 9 * C2; p.N = C3; p.K = M; p.A = n->a2; p.B = n->da3; p.ldb = C3; p.C = G[P_C3W]; p.ldc = C3;
    n->launches += launch_gemm<A_IM2COL_T, B_ROW, 128, C2>(p, n->partial_cap, n->n_sms, s);
    launch_colsum(n, n->da3, M, C3, G[P_C3B], s);
    p = GemmParams{}; p.batch = (int)B; p.partial = n->partial;
    p.M = M; p.N = C2; p.K = 9 * C3; p.A = n->da3; p.B = W[P_C3W]; p.C = n->da2; p.ldc = C2; p.mask_act = n->a2;
    n->launches += launch_gemm<A_IM2COL_FLIP, B_CONVW_T, 64, C3>(p, n->partial_cap, n->n_sms, s);
    // conv2
    p = GemmParams{}; p.batch = (int)B; p.partial = n->partial;
    p.M = 9 * C1; p.N = C2; p.K = M; p.A = n->a1; p.B = n->da2; p.ldb = C2; p.C = G[P_C2W]; p.ldc = C2;
    n->launches += launch_gemm<A_IM2COL_T, B_ROW, 64, C1>(p, n->partial_cap, n->n_sms, s);
    launch_colsum(n, n->da2, M, C2, G[P_C2B], s);
    p = GemmParams{}; p.batch = (int)B; p.partial = n->partial;
    p.M = M; p.N = C1; p.K = 9 * C2; p.A = n->da2; p.B = W[P_C2W]; p.C = n->da1; p.ldc = C1; p.mask_act = n->a1;
    n->launches += launch_gemm<A_IM2COL_FLIP, B_CONVW_T, 32, C2>(p, n->partial_cap, n->n_sms, s);
    // conv1 (no dgrad: the observation needs no gradient)
    p = GemmParams{}; p.batch = (int)B; p.partial = n->partial;
    p.M = 9 * CIN; p.N = C1; p.K = M; p.A = state; p.B = n->da1; p.ldb = C1; p.C = G[P_C1W]; p.ldc = C1;
    n->launches += launch_gemm<A_IM2COL_T, B_ROW, 32, CIN>(p, n->partial_cap, n->n_sms, s);
    launch_colsum(n, n->da1, M, C1, G[P_C1B], s);
}

static int td_backward_impl(mq_qnet* n, const float* state, const int64_t* action, const float* reward, const float* next_state,
                            const uint8_t* done, int64_t B, const mq_hparams* hp, const uint8_t* drop_online,
                            const uint8_t* drop_target, float* loss_out, void* stream, int part) {
    using namespace mq;
    MQ_REQUIRE(n && state && action && reward && next_state && done && hp && loss_out, "mq_qnet_td_backward: null argument");
    MQ_REQUIRE(B > 0 && B <= n->max_batch, "mq_qnet_td_backward: batch %lld outside 1..%lld", (long long)B, n->max_batch);
    MQ_REQUIRE(n->tl.g[0] && n->da3, "mq_qnet_td_backward: handle was created without gradient buffers");
    MQ_ON_DEVICE(n->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int hb = (int)((B * 32 + 255) / 256);
    // next_q = target_network(next_states).max(1)[0]   (dqn_agent.py:146-147)
    const bool bf16 = n->precision == 1;
    MQ_REQUIRE(!bf16 || B % 8 == 0, "mq_qnet_td_backward: the bf16 path needs a batch that is a multiple of 8 (got %lld)", (long long)B);
    if (part == 2) {                  // convolution layers of the backward only; part 1 has just run on the same batch
        if (bf16) {
            cudaError_t be = backward_bf16(n, state, B, drop_online, s, 2);
            if (be != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_qnet_td_backward (bf16 backward, part 2): %s", cudaGetErrorString(be));
        } else {
            backward_fp32(n, state, B, drop_online, s, 2);
        }
        MQ_CUDA(cudaGetLastError());
        return MQ_OK;
    }
    cudaError_t fe = cudaSuccess;
    if (bf16) fe = forward_net_bf16(n, 1, next_state, B, drop_target, s);
    else forward_net(n, n->tl.t, next_state, B, drop_target, s);
    qhead_kernel<<<hb, 256, 0, s>>>(n->h2, n->tl.t[P_F3W], n->tl.t[P_F3B], B, 3, nullptr, nullptr, n->maxq, nullptr, 0.f, 0, 0, 0, 1);
    // current_q = q_network(states).gather(1, actions)   (dqn_agent.py:143)
    if (bf16 && fe == cudaSuccess) fe = forward_net_bf16(n, 0, state, B, drop_online, s, true);
    else if (!bf16) forward_net(n, n->tl.p, state, B, drop_online, s);
    if (fe != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_qnet_td_backward (bf16 forward): %s", cudaGetErrorString(fe));
    qhead_kernel<<<hb, 256, 0, s>>>(n->h2, n->tl.p[P_F3W], n->tl.p[P_F3B], B, 2, nullptr, (const long long*)action, n->q_sa, nullptr, 0.f,
                                    0, 0, 0, 1);
    td_loss_kernel<<<1, 1024, 0, s>>>(n->q_sa, n->maxq, reward, done, (const long long*)action, B, hp->gamma, hp->huber, n->dq, loss_out);
    n->launches += 3;

    // ---- backward (loss.backward(), dqn_agent.py:154-155) ----
    if (bf16) {
        cudaError_t be = backward_bf16(n, state, B, drop_online, s, part);
        if (be != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_qnet_td_backward (bf16 backward): %s", cudaGetErrorString(be));
        return MQ_OK;
    }
    backward_fp32(n, state, B, drop_online, s, part);
    MQ_CUDA(cudaGetLastError());
    return MQ_OK;
}

extern "C" int mq_qnet_td_backward(mq_qnet* n, const float* state, const int64_t* action, const float* reward, const float* next_state,
                                   const uint8_t* done, int64_t B, const mq_hparams* hp, const uint8_t* drop_online,
                                   const uint8_t* drop_target, float* loss_out, void* stream) {
    return td_backward_impl(n, state, action, reward, next_state, done, B, hp, drop_online, drop_target, loss_out, stream, 0);
}

// The same step in two calls so that a data-parallel caller can overlap the gradient exchange with the backward: part 1 =
// both forwards, loss, backward of fc3 / fc2 / fc1 (99 % of the gradient: tensors 6..11 of the flat buffer are final when it
// returns); part 2 = backward of the three convolutions (tensors 0..5).  Same arguments for both parts.
extern "C" int mq_qnet_td_backward_part(mq_qnet* n, const float* state, const int64_t* action, const float* reward, const float* next_state,
                                        const uint8_t* done, int64_t B, const mq_hparams* hp, const uint8_t* drop_online,
                                        const uint8_t* drop_target, float* loss_out, int32_t part, void* stream) {
    MQ_REQUIRE(part == 1 || part == 2, "mq_qnet_td_backward_part: part must be 1 or 2");
    return td_backward_impl(n, state, action, reward, next_state, done, B, hp, drop_online, drop_target, loss_out, stream, part);
}

// Backward of the online network from an EXTERNAL dL/dQ (B x 5): what autograd does when a runner builds its own loss on
// agent.q_network(states) (train_qmix.py:92-110 mixes the Q-values of two agents before the loss).  The online forward is
// recomputed here (the workspace holds one set of activations and the caller may have run the target network since),
// then the same backward kernels as mq_qnet_td_backward fill the bound gradient buffers (overwriting them).
extern "C" int mq_qnet_backward(mq_qnet* n, const float* state, const float* dq, int64_t B, const uint8_t* drop_online, void* stream) {
    using namespace mq;
    MQ_REQUIRE(n && state && dq, "mq_qnet_backward: null argument");
    MQ_REQUIRE(B > 0 && B <= n->max_batch, "mq_qnet_backward: batch %lld outside 1..%lld", (long long)B, n->max_batch);
    MQ_REQUIRE(n->tl.g[0] && n->da3, "mq_qnet_backward: handle was created without gradient buffers");
    MQ_ON_DEVICE(n->device);
    cudaStream_t s = (cudaStream_t)stream;
    const bool bf16 = n->precision == 1;
    MQ_REQUIRE(!bf16 || B % 8 == 0, "mq_qnet_backward: the bf16 path needs a batch that is a multiple of 8 (got %lld)", (long long)B);
    if (bf16) {
        cudaError_t fe = forward_net_bf16(n, 0, state, B, drop_online, s, true);
        if (fe != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_qnet_backward (bf16 forward): %s", cudaGetErrorString(fe));
    } else {
        forward_net(n, n->tl.p, state, B, drop_online, s);
    }
    MQ_CUDA(cudaMemcpyAsync(n->dq, dq, sizeof(float) * (size_t)B * NA, cudaMemcpyDeviceToDevice, s));
    if (bf16) {
        cudaError_t be = backward_bf16(n, state, B, drop_online, s);
        if (be != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_qnet_backward (bf16 backward): %s", cudaGetErrorString(be));
    } else {
        backward_fp32(n, state, B, drop_online, s);
    }
    MQ_CUDA(cudaGetLastError());
    return MQ_OK;
}

extern "C" int mq_qnet_clip_adam(mq_qnet* n, const mq_hparams* hp, float grad_scale, float* gnorm_out, void* stream) {
    using namespace mq;
    MQ_REQUIRE(n && hp && hp->adam_step >= 1, "mq_qnet_clip_adam: bad argument");
    MQ_REQUIRE(n->tl.g[0] && n->tl.m[0] && n->tl.v[0], "mq_qnet_clip_adam: gradient / Adam buffers not bound");
    MQ_ON_DEVICE(n->device);
    cudaStream_t s = (cudaStream_t)stream;
    sqnorm_partial_kernel<<<(int)n->total_chunks, 256, 0, s>>>(n->tl, grad_scale, n->norm_partial);
    sqnorm_final_kernel<<<1, 1024, 0, s>>>(n->norm_partial, (int)n->total_chunks, n->gnorm);
    // scalars as torch.optim.Adam computes them (Python floats = double), then cast to fp32
    const double bc1 = 1.0 - std::pow((double)hp->beta1, (double)hp->adam_step);
    const double bc2 = 1.0 - std::pow((double)hp->beta2, (double)hp->adam_step);
    const float step_size = (float)((double)hp->lr / bc1);
    const float bc2_sqrt = (float)std::sqrt(bc2);
    clip_adam_kernel<<<(int)n->total_chunks, 256, 0, s>>>(n->tl, n->gnorm, grad_scale, hp->clip_norm, hp->beta1, hp->beta2, hp->adam_eps,
                                                          step_size, bc2_sqrt);
    if (gnorm_out) MQ_CUDA(cudaMemcpyAsync(gnorm_out, n->gnorm, sizeof(float), cudaMemcpyDeviceToDevice, s));
    n->launches += 3;
    n->w_dirty[0] = true;
    MQ_CUDA(cudaGetLastError());
    return MQ_OK;
}

extern "C" int mq_qnet_sync_target(mq_qnet* n, float tau, void* stream) {
    MQ_REQUIRE(n, "mq_qnet_sync_target: null handle");
    MQ_ON_DEVICE(n->device);
    mq::sync_target_kernel<<<(int)n->total_chunks, 256, 0, (cudaStream_t)stream>>>(n->tl, tau);
    n->launches += 1;
    n->w_dirty[1] = true;
    MQ_CUDA(cudaGetLastError());
    return MQ_OK;
}

extern "C" int mq_qnet_dropout_mask(uint8_t* mask, int64_t n, float p, uint64_t seed, uint64_t counter, void* stream) {
    MQ_REQUIRE(mask && n > 0 && p >= 0.f && p < 1.f, "mq_qnet_dropout_mask: bad argument");
    MQ_ON_DEVICE_OF(mask);
    const unsigned threshold = (unsigned)((double)p * 4294967296.0);
    const long long groups = (n + 3) / 4;
    mq::dropout_mask_kernel<<<(int)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mask, n, threshold, seed, counter);
    MQ_CUDA(cudaGetLastError());
    return MQ_OK;
}

extern "C" int mq_qnet_set_precision(mq_qnet* n, int32_t precision) {
    MQ_REQUIRE(n && (precision == 0 || precision == 1), "mq_qnet_set_precision: precision must be 0 (fp32) or 1 (bf16 tensor cores)");
    MQ_ON_DEVICE(n->device);
    if (precision == 1) {
        cudaError_t e = mq::alloc_bf16(n);
        if (e != cudaSuccess) return mq::fail(MQ_ERR_ALLOC, "mq_qnet_set_precision: bf16 workspace allocation failed: %s", cudaGetErrorString(e));
        if (!mq::tc::encode_fn()) return mq::fail(MQ_ERR_UNSUPPORTED, "mq_qnet_set_precision: cuTensorMapEncodeTiled is not available in this driver");
    }
    n->precision = precision;
    n->w_dirty[0] = n->w_dirty[1] = true;
    return MQ_OK;
}

extern "C" int mq_qnet_params_changed(mq_qnet* n) {
    MQ_REQUIRE(n, "mq_qnet_params_changed: null handle");
    n->w_dirty[0] = n->w_dirty[1] = true;
    return MQ_OK;
}
