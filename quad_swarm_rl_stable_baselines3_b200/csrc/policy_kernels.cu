// policy_kernels.cu -- fused two-tower policy forward for the quadrotor-swarm rollouts on sm_100a.
//
// What it replaces: the per-step forward of the reference's `ActorCriticPolicyCustomSeparateWeights`
// (swarm_rl/models/ActorCriticPolicyCustom.py:284-556) with one `QuadMultiEncoder` per tower
// (swarm_rl/models/quad_multi_model.py:250-354): self-observation MLP S -> 256 -> 256 (tanh), deep-sets neighbour encoder
// phi([self, nbr_j]) = (S + W) -> 256 -> 256 (tanh) averaged over the V visible neighbours (quad_multi_model.py:16-41), feed-forward
// 512 -> 512 (tanh), then the action-mean head (512 -> A) of the actor tower and the value head (512 -> 1) of the critic tower.
// This is the one dense contraction of the system: ~3.06 MFLOP per drone row, 1.6 TFLOP per 65536 x 8 step.
//
// Design (one CTA per SM, persistent over 128-row tiles, 128 threads = one thread per row = one TMEM lane):
//   * every GEMM is `tcgen05.mma.cta_group::1.kind::f16` (bf16 in, fp32 accumulate), M = 128 rows, N = 256, issued by one thread;
//   * ACTIVATIONS NEVER LEAVE TENSOR MEMORY: the accumulator (256 fp32 columns) is read back with `tcgen05.ld`, bias + tanh are
//     applied in registers, and the result is written with `tcgen05.st` as packed bf16 into a second TMEM region that the next
//     layer's MMA reads as its A operand (A from TMEM, "TS" form).  TMEM map (512 columns): [0,256) accumulator, [256,384) region
//     R1 (layer input / hidden / neighbour mean, K <= 256 bf16), [384,512) region R2 (self-encoder output, kept for the feed-forward);
//   * weights (B operand) are pre-packed on the device into the canonical no-swizzle K-major core-matrix layout (8 rows x 16 bytes),
//     one 32 KB image per (N = 256) x (K = 64) chunk, in the order the kernel consumes them, and streamed from L2 by 1-D bulk
//     copies (`cp.async.bulk` + mbarrier complete_tx) through a two-stage ring;
//   * the running sum over neighbours lives in shared memory as fp32 [256][128] (column-major: conflict-free per-row access);
//   * the 512 -> A / 512 -> 1 heads are folded into the feed-forward epilogue on the CUDA cores (4 FMA per element).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/quadpolicy.h"

namespace qp {

constexpr int H = 256;                      // hidden width of every encoder layer
constexpr int FF = 512;                     // feed-forward width
constexpr int TILE_M = 128;
constexpr int CHUNK_BYTES = 32 * 1024;      // one (N = 256) x (K = 64) bf16 weight image
constexpr int STAGES = 2;
constexpr int MAX_ACT = 8;

// fp32 side parameters of one tower, in this order
struct TowerParams {
    float b_self1[H], b_self2[H], b_nbr1[H], b_nbr2[H], b_ff[FF];
    float head_w[MAX_ACT * FF];             // [out][512], row-major like nn.Linear.weight
    float head_b[MAX_ACT];
};

struct Args {
    const float *obs; int n, stride, S, W, V, A;
    const __nv_bfloat16 *wchunks[2];        // per tower: chunk images in consumption order (5 + 5 per neighbour pass + 16)
    const TowerParams *params[2];
    float *mean, *value;                    // [n, A], [n]
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / bulk copy ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "QP_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra QP_DONE;\n"
        "bra QP_WAIT;\n"
        "QP_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- tcgen05 ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, M = 128, N = 256, K = 16, bf16 x bf16 -> fp32
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 B contiguous); LBO = distance between the two K-halves of one MMA,
// SBO = distance between 8-row groups (cute/atom/mma_traits_sm100.hpp, "LayoutType::INTERLEAVE ((8,n),2):((1,SBO),LBO)")
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                              // descriptor version (Blackwell)
    return d;                                                            // base offset 0, layout type 0 (no swizzle)
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tanh_fast(float x)
{
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));      // upper half <- first source
    return r;
}

constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);   // F32 acc, bf16 x bf16, K-major, N 256, M 128

// shared-memory map
struct Smem {
    uint8_t w[STAGES][CHUNK_BYTES];        // weight ring (1024-byte aligned by construction)
    float nbr_sum[H][TILE_M];              // running sum over neighbours, column-major
    uint64_t full[STAGES], empty[STAGES], acc_full;
    uint32_t tmem_base;
};

// the weight stream: thread 0 keeps one chunk in flight ahead of the one being consumed
struct Stream {
    const uint8_t *src;                    // next chunk image in global memory
    int issued, consumed;                  // chunk counters over the whole kernel (ring position and parity)
};

__device__ __forceinline__ void stream_issue(Smem &sm, Stream &st, const uint8_t *chunk, uint32_t bytes)
{
    const int s = st.issued % STAGES;
    mbar_wait(&sm.empty[s], ((st.issued / STAGES) & 1) ^ 1);            // the MMAs that read this slot have completed
    mbar_expect_tx(&sm.full[s], bytes);
    bulk_load(sm.w[s], chunk, bytes, &sm.full[s]);
    st.issued += 1;
}

// One GEMM: ACC[128 x 256] = A[128 x K] (TMEM columns a_col .. ) * W^T, W streamed as K / kc chunk images starting at `chunks`.
// Called by thread 0 only.  `next`: the first chunk of the following GEMM (prefetched while this GEMM's last chunk computes), or null.
__device__ __forceinline__ void gemm_issue(Smem &sm, Stream &st, uint32_t tmem, uint32_t acc_col, const uint32_t *a_cols, int nchunks, int kc,
                                           const uint8_t *chunks, const uint8_t *next, uint32_t next_bytes)
{
    const uint32_t bytes = (uint32_t)(H * kc * 2);
    const uint32_t sbo = (uint32_t)(kc / 8) * 128u;
    for (int c = 0; c < nchunks; ++c) {
        if (st.issued == st.consumed) stream_issue(sm, st, chunks + (size_t)c * CHUNK_BYTES, bytes);        // nothing in flight yet
        if (c + 1 < nchunks) stream_issue(sm, st, chunks + (size_t)(c + 1) * CHUNK_BYTES, bytes);            // one ahead
        else if (next != nullptr) stream_issue(sm, st, next, next_bytes);
        const int s = st.consumed % STAGES;
        mbar_wait(&sm.full[s], (st.consumed / STAGES) & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(sm.w[s]);
        for (int k = 0; k < kc / 16; ++k)
            mma_ts(tmem + acc_col, tmem + a_cols[c] + (uint32_t)(k * 8), smem_desc(base + (uint32_t)k * 256u, 128u, sbo), IDESC, (c | k) != 0);
        tc_commit(&sm.empty[s]);                                        // frees the slot when these MMAs have read it
        st.consumed += 1;
    }
    tc_commit(&sm.acc_full);
}

__global__ void __launch_bounds__(TILE_M, 1) policy_forward_kernel(const Args a)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        mbar_init(&sm.acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's 32 TMEM lanes
    constexpr uint32_t ACC = 0, R1 = 256, R2 = 384;
    Stream st; st.src = nullptr; st.issued = 0; st.consumed = 0;
    uint32_t acc_parity = 0;
    const int S = a.S, W = a.W, V = a.V, A = a.A;
    const int n_tiles = (a.n + TILE_M - 1) / TILE_M;

    // all threads: wait for the accumulator of the GEMM just issued
    auto wait_acc = [&]() { mbar_wait(&sm.acc_full, acc_parity); acc_parity ^= 1u; tc_fence_after(); };
    // all threads: activations written with tcgen05.st are ordered before the MMAs thread 0 issues next
    auto publish = [&]() { tmem_st_wait(); tc_fence_before(); __syncthreads(); tc_fence_after(); };

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row = tile * TILE_M + t;
        const bool live = row < a.n;
        const float *orow = a.obs + (size_t)(live ? row : 0) * a.stride;
        // the self observation, kept in registers as packed bf16 (reused by every neighbour pass)
        uint32_t xs[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float lo = (live && 2 * k < S) ? orow[2 * k] : 0.f, hi = (live && 2 * k + 1 < S) ? orow[2 * k + 1] : 0.f;
            xs[k] = pack_bf16(lo, hi);
        }
        for (int tower = 0; tower < 2; ++tower) {
            const uint8_t *wc = reinterpret_cast<const uint8_t *>(a.wchunks[tower]);
            const TowerParams *P = a.params[tower];
            // chunk images: [0] self L1 (K 32), [1..4] self L2, [5] nbr L1 (K 32), [6..9] nbr L2, [10..25] feed-forward (2 halves x 8)
            const uint8_t *c_self1 = wc, *c_self2 = wc + 1 * (size_t)CHUNK_BYTES, *c_nbr1 = wc + 5 * (size_t)CHUNK_BYTES,
                          *c_nbr2 = wc + 6 * (size_t)CHUNK_BYTES, *c_ff = wc + 10 * (size_t)CHUNK_BYTES;
            const uint32_t cols_l1[1] = { R1 }, cols_l2[4] = { R1, R1 + 32, R1 + 64, R1 + 96 };
            const uint32_t cols_ff[8] = { R2, R2 + 32, R2 + 64, R2 + 96, R1, R1 + 32, R1 + 64, R1 + 96 };

            // hidden-layer epilogue: ACC -> +bias -> tanh -> packed bf16 into TMEM region `dst` (the next layer's A operand)
            auto epilogue_to_tmem = [&](const float *bias, uint32_t dst) {
#pragma unroll 1
                for (int c0 = 0; c0 < H; c0 += 32) {
                    uint32_t v[32], u[16];
                    tmem_ld32(lane_addr + ACC + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        u[i] = pack_bf16(tanh_fast(__uint_as_float(v[2 * i]) + __ldg(bias + c0 + 2 * i)),
                                         tanh_fast(__uint_as_float(v[2 * i + 1]) + __ldg(bias + c0 + 2 * i + 1)));
                    tmem_st16(lane_addr + dst + (uint32_t)(c0 / 2), u);
                }
            };

            // ---- self encoder: S -> 256 -> 256
            tmem_st16(lane_addr + R1, xs);
            publish();
            if (t == 0) gemm_issue(sm, st, tmem, ACC, cols_l1, 1, 32, c_self1, c_self2, H * 64 * 2);
            wait_acc();
            epilogue_to_tmem(P->b_self1, R1);
            publish();
            if (t == 0) gemm_issue(sm, st, tmem, ACC, cols_l2, 4, 64, c_self2, c_nbr1, H * 32 * 2);
            wait_acc();
            epilogue_to_tmem(P->b_self2, R2);                          // stays in R2 until the feed-forward
            // ---- neighbour encoder (deep sets): mean_j tanh(W2 tanh(W1 [self, nbr_j] + b1) + b2)
            for (int j = 0; j < V; ++j) {
                uint32_t xn[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) xn[k] = xs[k];
                // append the neighbour's W values behind the S self values (S + W <= 32)
                for (int i = 0; i < W; ++i) {
                    const float v = live ? orow[S + j * W + i] : 0.f;
                    const int k = S + i;
                    const uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (q == (k >> 1)) xn[q] = (k & 1) ? ((xn[q] & 0x0000FFFFu) | (b << 16)) : ((xn[q] & 0xFFFF0000u) | b);
                }
                tmem_st16(lane_addr + R1, xn);
                publish();
                if (t == 0) gemm_issue(sm, st, tmem, ACC, cols_l1, 1, 32, c_nbr1, c_nbr2, H * 64 * 2);
                wait_acc();
                epilogue_to_tmem(P->b_nbr1, R1);
                publish();
                if (t == 0) gemm_issue(sm, st, tmem, ACC, cols_l2, 4, 64, c_nbr2, (j + 1 < V) ? c_nbr1 : c_ff, (j + 1 < V) ? H * 32 * 2 : H * 64 * 2);
                wait_acc();
#pragma unroll 1
                for (int c0 = 0; c0 < H; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(lane_addr + ACC + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float y = tanh_fast(__uint_as_float(v[i]) + __ldg(P->b_nbr2 + c0 + i));
                        sm.nbr_sum[c0 + i][t] = (j == 0) ? y : sm.nbr_sum[c0 + i][t] + y;
                    }
                }
                tc_fence_before();                                      // the tcgen05.ld above are ordered before the next GEMM overwrites ACC
            }
            // neighbour mean -> R1 (a tower without neighbours feeds zeros, like an absent encoder half)
            {
                const float inv = V > 0 ? 1.0f / (float)V : 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < H; c0 += 32) {
                    uint32_t u[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        u[i] = V > 0 ? pack_bf16(sm.nbr_sum[c0 + 2 * i][t] * inv, sm.nbr_sum[c0 + 2 * i + 1][t] * inv) : 0u;
                    tmem_st16(lane_addr + R1 + (uint32_t)(c0 / 2), u);
                }
            }
            publish();
            // ---- feed-forward 512 -> 512 (two halves of 256 columns) with the head folded into its epilogue
            const int n_out = tower == 0 ? A : 1;
            float head[MAX_ACT];
#pragma unroll
            for (int o = 0; o < MAX_ACT; ++o) head[o] = 0.f;
            for (int half = 0; half < 2; ++half) {
                const uint8_t *next = (half == 0) ? c_ff + 8 * (size_t)CHUNK_BYTES
                                                  : (tower == 0 ? reinterpret_cast<const uint8_t *>(a.wchunks[1])
                                                                : (tile + (int)gridDim.x < n_tiles ? reinterpret_cast<const uint8_t *>(a.wchunks[0]) : nullptr));
                if (t == 0) gemm_issue(sm, st, tmem, ACC, cols_ff, 8, 64, c_ff + (size_t)half * 8 * CHUNK_BYTES, next, half == 0 ? H * 64 * 2 : H * 32 * 2);
                wait_acc();
#pragma unroll 1
                for (int c0 = 0; c0 < H; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(lane_addr + ACC + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int c = half * H + c0 + i;
                        const float y = tanh_fast(__uint_as_float(v[i]) + __ldg(P->b_ff + c));
#pragma unroll
                        for (int o = 0; o < MAX_ACT; ++o)
                            if (o < n_out) head[o] = fmaf(y, __ldg(P->head_w + o * FF + c), head[o]);
                    }
                }
                tc_fence_before();
                __syncthreads();                                        // every row has read ACC before the next GEMM overwrites it
                tc_fence_after();
            }
            if (live) {
                if (tower == 0) { for (int o = 0; o < A; ++o) a.mean[(size_t)row * A + o] = head[o] + __ldg(P->head_b + o); }
                else a.value[row] = head[0] + __ldg(P->head_b);
            }
        }
    }
    // every issued chunk has been consumed (the prefetch pointers above never run past the last GEMM), TMEM can go
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- weight packing: nn.Linear weight [out, in] fp32 -> chunk images (bf16, canonical K-major core-matrix layout) ----------------
// chunk image of rows n0 .. n0+255 and inputs k0 .. k0+kc-1 of W (zero beyond `in`): byte offset of element (n, k) inside the image =
// (n / 8) * SBO + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2, SBO = (kc / 8) * 128.
__global__ void pack_chunk_kernel(const float *w, int out_dim, int in_dim, int n0, int k0, int kc, __nv_bfloat16 *img)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * kc) return;
    const int n = idx / kc, k = idx % kc;
    const int gn = n0 + n, gk = k0 + k;
    const float v = (gn < out_dim && gk < in_dim) ? w[(size_t)gn * in_dim + gk] : 0.f;
    const size_t off = (size_t)(n / 8) * ((size_t)(kc / 8) * 128) + (size_t)(k / 8) * 128 + (size_t)(n % 8) * 16 + (size_t)(k % 8) * 2;
    img[off / 2] = __float2bfloat16_rn(v);
}

}  // namespace qp

using namespace qp;

struct qp_policy {
    qp_config cfg;
    int device, sms;
    __nv_bfloat16 *wchunks[2];
    TowerParams *params[2];
    long long launches;
    std::string err;
};

static thread_local std::string g_qp_err;
static int qp_fail(qp_policy *p, int code, const std::string &m) { if (p) p->err = m; else g_qp_err = m; return code; }
#define QP_CUDA(p, call)                                                                                          \
    do {                                                                                                          \
        cudaError_t _r = (call);                                                                                  \
        if (_r != cudaSuccess) return qp_fail(p, QP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_r)); \
    } while (0)

constexpr int CHUNKS_PER_TOWER = 26;

extern "C" {

const char *qp_last_error(const qp_policy *p) { return p ? p->err.c_str() : g_qp_err.c_str(); }
size_t qp_config_size(void) { return sizeof(qp_config); }

int qp_create(const qp_config *cfg, int device, qp_policy **out)
{
    if (!cfg || !out) return qp_fail(nullptr, QP_ERR_NULL, "qp_create: null argument");
    *out = nullptr;
    if (cfg->api_version != QP_API_VERSION) return qp_fail(nullptr, QP_ERR_BAD_CONFIG, "qp_create: api_version mismatch");
    if (cfg->hidden != H) return qp_fail(nullptr, QP_ERR_BAD_CONFIG, "qp_create: only hidden = 256 is built");
    if (cfg->self_dim < 1 || cfg->nbr_dim < 0 || cfg->num_nbr < 0 || cfg->self_dim + cfg->nbr_dim > 32)
        return qp_fail(nullptr, QP_ERR_BAD_CONFIG, "qp_create: self_dim + nbr_dim must be <= 32");
    if (cfg->act_dim < 1 || cfg->act_dim > MAX_ACT) return qp_fail(nullptr, QP_ERR_BAD_CONFIG, "qp_create: act_dim out of [1, 8]");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return qp_fail(nullptr, QP_ERR_CUDA, "qp_create: no such CUDA device");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    qp_policy *p = new qp_policy();
    p->cfg = *cfg; p->device = device; p->launches = 0;
    cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, device);
    cudaError_t r = cudaSuccess;
    for (int t = 0; t < 2 && r == cudaSuccess; ++t) {
        r = cudaMalloc(&p->wchunks[t], (size_t)CHUNKS_PER_TOWER * CHUNK_BYTES);
        if (r == cudaSuccess) r = cudaMemset(p->wchunks[t], 0, (size_t)CHUNKS_PER_TOWER * CHUNK_BYTES);
        if (r == cudaSuccess) r = cudaMalloc(&p->params[t], sizeof(TowerParams));
        if (r == cudaSuccess) r = cudaMemset(p->params[t], 0, sizeof(TowerParams));
    }
    if (r == cudaSuccess) r = cudaFuncSetAttribute(policy_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024);
    cudaSetDevice(prev);
    if (r != cudaSuccess) { delete p; return qp_fail(nullptr, QP_ERR_CUDA, std::string("qp_create: ") + cudaGetErrorString(r)); }
    *out = p;
    return QP_OK;
}

int qp_destroy(qp_policy *p)
{
    if (!p) return QP_ERR_NULL;
    for (int t = 0; t < 2; ++t) { cudaFree(p->wchunks[t]); cudaFree(p->params[t]); }
    delete p;
    return QP_OK;
}

int64_t qp_launch_count(const qp_policy *p) { return p ? p->launches : 0; }

int qp_set_weights(qp_policy *p, int tower, const qp_tower_weights *w, void *stream)
{
    if (!p || !w) return qp_fail(p, QP_ERR_NULL, "qp_set_weights: null argument");
    if (tower < 0 || tower > 1) return qp_fail(p, QP_ERR_BAD_CONFIG, "qp_set_weights: tower must be 0 (actor) or 1 (critic)");
    cudaStream_t s = (cudaStream_t)stream;
    const int S = p->cfg.self_dim, SW = p->cfg.self_dim + p->cfg.nbr_dim;
    const int n_out = tower == 0 ? p->cfg.act_dim : 1;
    __nv_bfloat16 *base = p->wchunks[tower];
    auto img = [&](int c) { return base + (size_t)c * (CHUNK_BYTES / 2); };
    auto pack = [&](const float *src, int out_dim, int in_dim, int n0, int k0, int kc, int chunk) {
        pack_chunk_kernel<<<(H * kc + 255) / 256, 256, 0, s>>>(src, out_dim, in_dim, n0, k0, kc, img(chunk));
    };
    pack(w->self_w1, H, S, 0, 0, 32, 0);
    for (int c = 0; c < 4; ++c) pack(w->self_w2, H, H, 0, 64 * c, 64, 1 + c);
    if (p->cfg.num_nbr > 0) {
        pack(w->nbr_w1, H, SW, 0, 0, 32, 5);
        for (int c = 0; c < 4; ++c) pack(w->nbr_w2, H, H, 0, 64 * c, 64, 6 + c);
    }
    for (int half = 0; half < 2; ++half)
        for (int c = 0; c < 8; ++c) pack(w->ff_w, FF, FF, 256 * half, 64 * c, 64, 10 + half * 8 + c);
    TowerParams *P = p->params[tower];
    const cudaMemcpyKind dd = cudaMemcpyDeviceToDevice;
    QP_CUDA(p, cudaMemcpyAsync(P->b_self1, w->self_b1, H * sizeof(float), dd, s));
    QP_CUDA(p, cudaMemcpyAsync(P->b_self2, w->self_b2, H * sizeof(float), dd, s));
    if (p->cfg.num_nbr > 0) {
        QP_CUDA(p, cudaMemcpyAsync(P->b_nbr1, w->nbr_b1, H * sizeof(float), dd, s));
        QP_CUDA(p, cudaMemcpyAsync(P->b_nbr2, w->nbr_b2, H * sizeof(float), dd, s));
    }
    QP_CUDA(p, cudaMemcpyAsync(P->b_ff, w->ff_b, FF * sizeof(float), dd, s));
    QP_CUDA(p, cudaMemcpyAsync(P->head_w, w->head_w, (size_t)n_out * FF * sizeof(float), dd, s));
    QP_CUDA(p, cudaMemcpyAsync(P->head_b, w->head_b, (size_t)n_out * sizeof(float), dd, s));
    QP_CUDA(p, cudaGetLastError());
    return QP_OK;
}

int qp_forward(qp_policy *p, const float *obs, int n, int obs_stride, float *mean, float *value, void *stream)
{
    if (!p || !obs || !mean || !value) return qp_fail(p, QP_ERR_NULL, "qp_forward: null argument");
    if (n < 1 || obs_stride < p->cfg.self_dim + p->cfg.nbr_dim * p->cfg.num_nbr) return qp_fail(p, QP_ERR_BAD_CONFIG, "qp_forward: bad n / obs_stride");
    Args a;
    a.obs = obs; a.n = n; a.stride = obs_stride; a.S = p->cfg.self_dim; a.W = p->cfg.nbr_dim; a.V = p->cfg.num_nbr; a.A = p->cfg.act_dim;
    for (int t = 0; t < 2; ++t) { a.wchunks[t] = p->wchunks[t]; a.params[t] = p->params[t]; }
    a.mean = mean; a.value = value;
    const int tiles = (n + TILE_M - 1) / TILE_M;
    const int grid = tiles < p->sms ? tiles : p->sms;
    policy_forward_kernel<<<grid, TILE_M, sizeof(Smem) + 1024, (cudaStream_t)stream>>>(a);
    p->launches += 1;
    QP_CUDA(p, cudaGetLastError());
    return QP_OK;
}

}  // extern "C"
