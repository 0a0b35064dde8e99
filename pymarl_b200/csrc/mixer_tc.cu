// bf16 tensor-core tier, QMIX mixer backward on tile images (modules/mixers/qmix.py:28-47 differentiated by hand;
// the reference gets it from torch autograd at learners/q_learner.py:101).
//
//   mix_bwd_img_kernel : one warp per (b, t) row.  Reads the 64-column blocks of the raw image row (128 contiguous
//                        bytes per block, coalesced), recomputes pre = sum_n q_n |w1_n| + b1, and overwrites the row with
//                        d_raw:  d_w1[n] = sign(w1[n]) q_n dpre,  d_b1 = dpre,  d_wf = sign(wf) g hidden,
//                        d_v0 = [v0 > 0] g V2,   dpre = g |wf| ELU'(pre);   d_q[n] = |w1[n]| . dpre;  V.2 gradients.
//   mix_dw_tc_kernel   : dW_cat | db_cat = d_raw^T . [state | 1]  (reduction over the B*T rows) on tcgen05.  Both
//                        operands are bulk-copied tile images read MN-major (no conversion): a CTA owns a 256 x 256
//                        output tile (two 128 x 256 fp32 accumulators = all 512 TMEM columns) and a slice of the rows;
//                        64 rows per pipeline stage, 3 stages.  Slices are summed in a fixed order (deterministic).
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "mixer_tc.cuh"

namespace pmb {
namespace tc {

__device__ __forceinline__ float mt_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float mt_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float mt_sgn(float x) { return (float)((x > 0.f) - (x < 0.f)); }

struct MixBwdParams {
    uint8_t* raw_img;            // in: raw, out: d_raw
    const float* agent_qs;       // [B*(T-1)][N]
    const float* g;              // [B*(T-1)]   dL/dQ_tot (un-normalised)
    const float* v2_w;           // [32]
    float* d_qs;                 // [B*(T-1)][N]
    float* v2_partial;           // [grid][33]
    int64_t BT, rows_total;      // B*T, row tiles * 128
    int T, N, n_cblk;
};

constexpr int MB_WARPS = 8;

template <int MAXCB>
__global__ void __launch_bounds__(32 * MB_WARPS) mix_bwd_img_kernel(MixBwdParams P) {
    __shared__ float red[MB_WARPS][33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * MB_WARPS;          // multiple of 8: (row & 7) is fixed per warp
    const int N = P.N;
    float dv2w0 = 0.f, dv2w1 = 0.f, dv2b = 0.f;
    // lane -> logical position inside a 64-column block (the 128-byte swizzle XORs the chunk index with row & 7)
    const int j = (lane >> 2) ^ wib;                                // logical 16-byte chunk (row & 7 == wib)
    const int h = j >> 2;                                           // group parity: this lane sees group 2*cb + h
    const int e0 = (j & 3) * 8 + (lane & 3) * 2;                    // its two hypernet-embed indices e0, e0 + 1
    const float v2w0 = __ldg(P.v2_w + e0), v2w1 = __ldg(P.v2_w + e0 + 1);

    for (int64_t m = (int64_t)blockIdx.x * MB_WARPS + wib; m < P.rows_total; m += n_warps) {
        uint8_t* base = P.raw_img + (m >> 7) * (int64_t)P.n_cblk * 16384 + (m & 127) * 128 + lane * 4;
        int64_t mq = -1;
        if (m < P.BT) {
            const int64_t b = m / P.T;
            const int t = (int)(m - b * P.T);
            if (t < P.T - 1) mq = b * (P.T - 1) + t;
        }
        if (mq < 0) {                                               // no online-mixer row here: d_raw = 0
            for (int cb = 0; cb < P.n_cblk; ++cb) *reinterpret_cast<uint32_t*>(base + (int64_t)cb * 16384) = 0u;
            continue;
        }
        const float gm = __ldg(P.g + mq);
        const float* qs = P.agent_qs + mq * N;
        // the whole row goes to registers first (all loads in flight), q_n of this lane's groups alongside
        uint32_t wv[MAXCB];
        float qv[MAXCB];
#pragma unroll
        for (int cb = 0; cb < MAXCB; ++cb) {
            wv[cb] = 0u; qv[cb] = 0.f;
            if (cb < P.n_cblk) {
                wv[cb] = *reinterpret_cast<const uint32_t*>(base + (int64_t)cb * 16384);
                if (2 * cb + h < N) qv[cb] = __ldg(qs + 2 * cb + h);
            }
        }
        // pre-activation of the mixing layer and the three special groups
        float acc0 = 0.f, acc1 = 0.f;
        uint32_t w_b1 = 0u, w_wf = 0u, w_v0 = 0u;
#pragma unroll
        for (int cb = 0; cb < MAXCB; ++cb) {
            const int grp = 2 * cb + h;
            if (grp < N) {
                acc0 = fmaf(qv[cb], fabsf(mt_lo(wv[cb])), acc0);
                acc1 = fmaf(qv[cb], fabsf(mt_hi(wv[cb])), acc1);
            } else if (grp == N) w_b1 = wv[cb];
            else if (grp == N + 1) w_wf = wv[cb];
            else if (grp == N + 2) w_v0 = wv[cb];
        }
        // the partner lane (lane ^ 16) holds the same e0 for the other group parity
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 16);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 16);
        const uint32_t o_b1 = __shfl_xor_sync(0xffffffffu, w_b1, 16), o_wf = __shfl_xor_sync(0xffffffffu, w_wf, 16),
                       o_v0 = __shfl_xor_sync(0xffffffffu, w_v0, 16);
        if (h != (N & 1)) w_b1 = o_b1;
        if (h != ((N + 1) & 1)) w_wf = o_wf;
        if (h != ((N + 2) & 1)) w_v0 = o_v0;
        const float pre0 = acc0 + mt_lo(w_b1), pre1 = acc1 + mt_hi(w_b1);
        const float hid0 = pre0 > 0.f ? pre0 : expm1f(pre0), hid1 = pre1 > 0.f ? pre1 : expm1f(pre1);
        const float wf0 = mt_lo(w_wf), wf1 = mt_hi(w_wf), v00 = mt_lo(w_v0), v01 = mt_hi(w_v0);
        const float dp0 = gm * fabsf(wf0) * (pre0 > 0.f ? 1.f : __expf(pre0));
        const float dp1 = gm * fabsf(wf1) * (pre1 > 0.f ? 1.f : __expf(pre1));
        const uint32_t d_wf = pack_bf16x2(mt_sgn(wf0) * gm * hid0, mt_sgn(wf1) * gm * hid1);
        const uint32_t d_b1 = pack_bf16x2(dp0, dp1);
        const uint32_t d_v0 = pack_bf16x2(v00 > 0.f ? gm * v2w0 : 0.f, v01 > 0.f ? gm * v2w1 : 0.f);
        if (h == 0) {                                               // one of the two lanes that hold e0 accumulates
            dv2w0 = fmaf(gm, fmaxf(v00, 0.f), dv2w0);
            dv2w1 = fmaf(gm, fmaxf(v01, 0.f), dv2w1);
        }
        if (lane == 0) dv2b += gm;
        // d_raw in place, d_q
#pragma unroll
        for (int cb = 0; cb < MAXCB; ++cb) {
            if (cb < P.n_cblk) {
                const uint32_t w = wv[cb];
                const int grp = 2 * cb + h;
                uint32_t out = 0u;
                float part = 0.f;
                if (grp < N) {
                    const float a = mt_lo(w), b = mt_hi(w);
                    out = pack_bf16x2(mt_sgn(a) * qv[cb] * dp0, mt_sgn(b) * qv[cb] * dp1);
                    part = fmaf(fabsf(a), dp0, fabsf(b) * dp1);
                } else if (grp == N) out = d_b1;
                else if (grp == N + 1) out = d_wf;
                else if (grp == N + 2) out = d_v0;
                *reinterpret_cast<uint32_t*>(base + (int64_t)cb * 16384) = out;
                // sum over the 16 lanes of this parity (lane bits 0..3)
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                part += __shfl_xor_sync(0xffffffffu, part, 2);
                part += __shfl_xor_sync(0xffffffffu, part, 4);
                part += __shfl_xor_sync(0xffffffffu, part, 8);
                if ((lane & 15) == 0 && grp < N) P.d_qs[mq * N + grp] = part;
            }
        }
    }
    // V.2 gradients: fixed lane -> e mapping per warp, fixed warp order -> deterministic partials
    for (int i = lane; i < 33; i += 32) red[wib][i] = 0.f;
    __syncwarp();
    if (h == 0) { red[wib][e0] = dv2w0; red[wib][e0 + 1] = dv2w1; }
    if (lane == 0) red[wib][32] = dv2b;
    __syncthreads();
    if (threadIdx.x < 33) {
        float s = 0.f;
        for (int w = 0; w < MB_WARPS; ++w) s += red[w][threadIdx.x];
        P.v2_partial[(int64_t)blockIdx.x * 33 + threadIdx.x] = s;
    }
}

// Lean variant for up to 16 column blocks (N <= 29), ~3x fewer instructions per row than the generic kernel above
// (which was issue-bound at 1370 warp instructions per row): sign(w) is applied to both bf16 halves of a word with
// one XOR on the packed product, group membership tests are hoisted into two per-lane bounds, and the 16 per-lane
// d_q partials are reduced with a 4-stage reduce-scatter (15 shuffles instead of 60) that leaves one group sum per
// lane for a single coalesced store.
//
// NT > 0: the agent count as a compile-time constant.  The ncu source page of the run-time form showed 1107 warp
// instructions per row, a third of them ISETP / SEL: every `cb < n_cblk`, `cb < n_w1`, `cb == cb_b1` test of the unrolled
// 16-block loops was re-evaluated per row (the kernel ran at IPC 2.7 and 34 % of the DRAM bandwidth: issue-bound).  With N
// known all of them fold (only block N / 2 keeps a per-lane parity test).  NT = 0 keeps the run-time form.
template <int NT>
__global__ void __launch_bounds__(32 * MB_WARPS, 3) mix_bwd_img16_kernel(MixBwdParams P) {
    __shared__ float red[MB_WARPS][33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * MB_WARPS;          // multiple of 8: (row & 7) is fixed per warp
    const int N = NT ? NT : P.N;
    const int NCB = NT ? (NT + 4) / 2 : P.n_cblk;                   // == tc_mix_cblks
    float dv2w0 = 0.f, dv2w1 = 0.f, dv2b = 0.f;
    const int j = (lane >> 2) ^ wib;                                // logical 16-byte chunk (row & 7 == wib)
    const int h = j >> 2;                                           // group parity: this lane sees group 2*cb + h
    const int e0 = (j & 3) * 8 + (lane & 3) * 2;
    const float v2w0 = __ldg(P.v2_w + e0), v2w1 = __ldg(P.v2_w + e0 + 1);
    const int n_w1 = (N - h + 1) >> 1;                              // cb < n_w1  <=>  group 2*cb + h is a w1 group
    // the same test in a form that folds for a constant N: blocks below N / 2 hold w1 groups for both parities, block
    // N / 2 only for an odd N and the even parity
    auto is_w1 = [&](int cb) -> bool {
        return NT ? (cb < NT / 2 || (cb == NT / 2 && (NT & 1) != 0 && h == 0)) : cb < n_w1;
    };
    // block index / need-partner flags of the three special groups for this lane
    const int cb_b1 = N >> 1, cb_wf = (N + 1) >> 1, cb_v0 = (N + 2) >> 1;
    const bool sw_b1 = h != (N & 1), sw_wf = h != ((N + 1) & 1), sw_v0 = h != ((N + 2) & 1);
    const int my_grp = 2 * (lane & 15) + h;                         // the group whose d_q this lane ends up holding

    for (int64_t m = (int64_t)blockIdx.x * MB_WARPS + wib; m < P.rows_total; m += n_warps) {
        uint8_t* base = P.raw_img + (m >> 7) * (int64_t)NCB * 16384 + (m & 127) * 128 + lane * 4;
        int64_t mq = -1;
        if (m < P.BT) {
            const uint32_t b = (uint32_t)m / (uint32_t)P.T;          // B*T < 2^31 (checked by the launcher)
            const int t = (int)((uint32_t)m - b * (uint32_t)P.T);
            if (t < P.T - 1) mq = (int64_t)b * (P.T - 1) + t;
        }
        if (mq < 0) {                                               // no online-mixer row here: d_raw = 0
            for (int cb = 0; cb < NCB; ++cb) *reinterpret_cast<uint32_t*>(base + (int64_t)cb * 16384) = 0u;
            continue;
        }
        const float gm = __ldg(P.g + mq);
        const float* qs = P.agent_qs + mq * N + h;
        uint32_t wv[16];
        float qv[16];
#pragma unroll
        for (int cb = 0; cb < 16; ++cb) {
            wv[cb] = 0u; qv[cb] = 0.f;
            if (cb < NCB) wv[cb] = *reinterpret_cast<const uint32_t*>(base + (int64_t)cb * 16384);
            if (is_w1(cb)) qv[cb] = __ldg(qs + 2 * cb);
        }
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int cb = 0; cb < 16; ++cb) {                           // qv = 0 outside the w1 groups
            acc0 = fmaf(qv[cb], fabsf(mt_lo(wv[cb])), acc0);
            acc1 = fmaf(qv[cb], fabsf(mt_hi(wv[cb])), acc1);
        }
        uint32_t w_b1 = 0u, w_wf = 0u, w_v0 = 0u;
#pragma unroll
        for (int cb = 0; cb < 16; ++cb) {
            w_b1 = cb == cb_b1 ? wv[cb] : w_b1;
            w_wf = cb == cb_wf ? wv[cb] : w_wf;
            w_v0 = cb == cb_v0 ? wv[cb] : w_v0;
        }
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 16);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 16);
        const uint32_t o_b1 = __shfl_xor_sync(0xffffffffu, w_b1, 16), o_wf = __shfl_xor_sync(0xffffffffu, w_wf, 16),
                       o_v0 = __shfl_xor_sync(0xffffffffu, w_v0, 16);
        w_b1 = sw_b1 ? o_b1 : w_b1; w_wf = sw_wf ? o_wf : w_wf; w_v0 = sw_v0 ? o_v0 : w_v0;
        const float pre0 = acc0 + mt_lo(w_b1), pre1 = acc1 + mt_hi(w_b1);
        const float hid0 = pre0 > 0.f ? pre0 : expm1f(pre0), hid1 = pre1 > 0.f ? pre1 : expm1f(pre1);
        const float wf0 = mt_lo(w_wf), wf1 = mt_hi(w_wf), v00 = mt_lo(w_v0), v01 = mt_hi(w_v0);
        const float dp0 = gm * fabsf(wf0) * (pre0 > 0.f ? 1.f : __expf(pre0));
        const float dp1 = gm * fabsf(wf1) * (pre1 > 0.f ? 1.f : __expf(pre1));
        const uint32_t d_wf = pack_bf16x2(mt_sgn(wf0) * gm * hid0, mt_sgn(wf1) * gm * hid1);
        const uint32_t d_b1 = pack_bf16x2(dp0, dp1);
        const uint32_t d_v0 = pack_bf16x2(v00 > 0.f ? gm * v2w0 : 0.f, v01 > 0.f ? gm * v2w1 : 0.f);
        if (h == 0) {
            dv2w0 = fmaf(gm, fmaxf(v00, 0.f), dv2w0);
            dv2w1 = fmaf(gm, fmaxf(v01, 0.f), dv2w1);
        }
        if (lane == 0) dv2b += gm;
        float part[16];
#pragma unroll
        for (int cb = 0; cb < 16; ++cb) {
            const uint32_t w = wv[cb];
            // d_w1 = sign(w) q dpre : sign bits XORed into the packed product, exact zeros of w (sign(0) = 0) masked out
            const uint32_t prod = pack_bf16x2(qv[cb] * dp0, qv[cb] * dp1);
            const uint32_t mag = w & 0x7fff7fffu;
            const uint32_t nz = (((mag + 0x7fff7fffu) | mag) & 0x80008000u) >> 15;      // 1 per non-zero half
            uint32_t out = (prod ^ (w & 0x80008000u)) & (nz * 0xffffu);
            part[cb] = fmaf(fabsf(mt_lo(w)), dp0, fabsf(mt_hi(w)) * dp1);               // garbage for non-w1 groups: not stored
            out = is_w1(cb) ? out : 0u;
            out = cb == cb_b1 && !sw_b1 ? d_b1 : out;
            out = cb == cb_wf && !sw_wf ? d_wf : out;
            out = cb == cb_v0 && !sw_v0 ? d_v0 : out;
            if (cb < NCB) *reinterpret_cast<uint32_t*>(base + (int64_t)cb * 16384) = out;
        }
        // reduce-scatter over the 16 lanes of a parity: after the stage with distance d a lane keeps the half of
        // its values selected by (lane & d); lane l ends with the sum for block l & 15
#pragma unroll
        for (int d = 8; d >= 1; d >>= 1) {
            const bool up = (lane & d) != 0;
#pragma unroll
            for (int i = 0; i < d; ++i) {
                const float send = up ? part[i] : part[i + d];
                const float keep = up ? part[i + d] : part[i];
                part[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
            }
        }
        if (my_grp < N) P.d_qs[mq * N + my_grp] = part[0];
    }
    for (int i = lane; i < 33; i += 32) red[wib][i] = 0.f;
    __syncwarp();
    if (h == 0) { red[wib][e0] = dv2w0; red[wib][e0 + 1] = dv2w1; }
    if (lane == 0) red[wib][32] = dv2b;
    __syncthreads();
    if (threadIdx.x < 33) {
        float s = 0.f;
        for (int w = 0; w < MB_WARPS; ++w) s += red[w][threadIdx.x];
        P.v2_partial[(int64_t)blockIdx.x * 33 + threadIdx.x] = s;
    }
}

// one block per output (32 V.2 weights + the V.2 bias); fixed summation order: deterministic
__global__ void __launch_bounds__(256) mix_v2_reduce_kernel(const float* __restrict__ partial, int n_blocks, float* __restrict__ dv2_w,
                                                            float* __restrict__ dv2_b) {
    __shared__ float red[8];
    const int i = blockIdx.x;
    float s = 0.f;
    for (int b = threadIdx.x; b < n_blocks; b += 256) s += partial[(int64_t)b * 33 + i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k];
        if (i < 32) dv2_w[i] = t; else dv2_b[0] = t;
    }
}

// ------------------------------------------------------------------------------------------
// hypernet weight-gradient GEMM
// ------------------------------------------------------------------------------------------
namespace md {
constexpr int HALF = 8192;                        // 64 rows x 128 B : half a tile image
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 8 * HALF;             // d_raw halves 0..3 | state halves 0..3
constexpr int BARS = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = 1024 + BARS + 128;
constexpr int THREADS = 192;                      // warps 0-3 epilogue, 4 MMA issuer, 5 loader
}  // namespace md

struct MixDwParams {
    const uint8_t* raw_img;          // d_raw images [row tile][n_cblk][16 KB]
    const uint8_t* state_img;        // [row tile][n_chunks][16 KB]
    float* partial;                  // [slices][n_ct*256][ldk]
    int n_cblk, n_chunks, n_ct, n_kt, ldk;
    int64_t n_halves, halves_per_slice;
};

__global__ void __launch_bounds__(md::THREADS, 1) mix_dw_tc_kernel(MixDwParams P) {
    using namespace md;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BARS);
    uint64_t* full = bars;               // [STAGES]
    uint64_t* empty = bars + STAGES;     // [STAGES]
    uint64_t* done = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int n_tile = P.n_ct * P.n_kt;
    const int tile = blockIdx.x % n_tile, slice = blockIdx.x / n_tile;
    const int ct = tile / P.n_kt, kt = tile - ct * P.n_kt;
    const int nD = P.n_cblk - 4 * ct < 4 ? P.n_cblk - 4 * ct : 4;          // 64-column blocks of d_raw in this tile
    const int nA = P.n_chunks - 4 * kt < 4 ? P.n_chunks - 4 * kt : 4;      // 64-column chunks of state
    const int ncols = nA * 64;
    const int64_t beg = (int64_t)slice * P.halves_per_slice;
    const int64_t end = beg + P.halves_per_slice < P.n_halves ? beg + P.halves_per_slice : P.n_halves;
    const int64_t n_my = end > beg ? end - beg : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    // d_raw halves that do not exist in this tile stay zero in every stage
    for (int i = threadIdx.x; i < STAGES * (4 - nD) * (HALF / 16); i += THREADS) {
        const int per = (4 - nD) * (HALF / 16);
        const int s = i / per, q = i - s * per;
        reinterpret_cast<uint4*>(smem + s * STAGE_BYTES + nD * HALF)[q] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 5) {
        if (lane == 0) {
            for (int64_t i = 0; i < n_my; ++i) {
                const int s = (int)(i % STAGES);
                const int64_t hh = beg + i, rt = hh >> 1;
                const int64_t hoff = (hh & 1) * HALF;
                mbar_wait(&empty[s], (uint32_t)(((i / STAGES) & 1) ^ 1));
                mbar_arrive_expect_tx(&full[s], (uint32_t)((nD + nA) * HALF));
                uint8_t* st = smem + s * STAGE_BYTES;
                for (int q = 0; q < nD; ++q)
                    bulk_copy_g2s(st + q * HALF, P.raw_img + (rt * P.n_cblk + 4 * ct + q) * 16384 + hoff, HALF, &full[s]);
                for (int q = 0; q < nA; ++q)
                    bulk_copy_g2s(st + (4 + q) * HALF, P.state_img + (rt * P.n_chunks + 4 * kt + q) * 16384 + hoff, HALF,
                                  &full[s]);
            }
        }
    } else if (warp == 4) {
        if (lane == 0 && n_my > 0) {
            const uint32_t idesc = umma_idesc_bf16(128, ncols, 1, 1);
            for (int64_t i = 0; i < n_my; ++i) {
                const int s = (int)(i % STAGES);
                mbar_wait(&full[s], (uint32_t)((i / STAGES) & 1));
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {                   // 64 rows = 4 x (K = 16)
                    const uint32_t acc = (i | kk) != 0;
                    const uint64_t bdesc = umma_desc_sw128(st + 4 * HALF + kk * 2048, HALF, 1024);
                    umma_bf16(tmem_base, umma_desc_sw128(st + kk * 2048, HALF, 1024), bdesc, idesc, acc);
                    if (nD > 2)
                        umma_bf16(tmem_base + 256, umma_desc_sw128(st + 2 * HALF + kk * 2048, HALF, 1024), bdesc, idesc, acc);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(done);
        }
    } else {
        // epilogue: accumulator row = d_raw column, accumulator column = state column
        float* out = P.partial + ((int64_t)slice * P.n_ct * 256 + ct * 256 + warp * 32 + lane) * P.ldk + kt * 256;
        const int n_acc = nD > 2 ? 2 : 1;
        if (n_my > 0) {
            mbar_wait(done, 0);
            tc_fence_after();
            const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int a = 0; a < n_acc; ++a)
                for (int g = 0; g < ncols; g += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tl + a * 256 + g, v);
                    tmem_wait_ld();
                    float4* o = reinterpret_cast<float4*>(out + (int64_t)a * 128 * P.ldk + g);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        o[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                           __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                }
        } else {
            for (int a = 0; a < n_acc; ++a)
                for (int g = 0; g < ncols; ++g) out[(int64_t)a * 128 * P.ldk + g] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// packed column c = grp*32 + e  (groups [w1 .. | b1 | w_final | v0])  ->  flat row [w1 .. | w_final | b1 | v0]
__global__ void __launch_bounds__(256)
mix_dw_reduce_kernel(const float* __restrict__ partial, int slices, int c_rows, int ldk, int N, int S,
                     float* __restrict__ gw_cat, float* __restrict__ gb_cat) {
    const int S1 = S + 1;
    const int64_t total = (int64_t)(N + 3) * 32 * S1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i / S1), k = (int)(i - (int64_t)c * S1);
    float s = 0.f;
    for (int sl = 0; sl < slices; ++sl) s += partial[((int64_t)sl * c_rows + c) * ldk + k];
    const int grp = c >> 5, e = c & 31;
    const int fgrp = grp < N ? grp : (grp == N ? N + 1 : (grp == N + 1 ? N : N + 2));
    if (k < S) gw_cat[((int64_t)fgrp * 32 + e) * S + k] = s;
    else gb_cat[fgrp * 32 + e] = s;
}

}  // namespace tc

namespace {
struct MixDwPlan { int n_ct, n_kt, slices, ldk; int64_t n_halves, halves_per_slice; };
// ctas_avail: SMs the weight-gradient GEMM may take (0 = all of them); it runs next to the agent's BPTT kernel on the
// side stream when that kernel leaves SMs idle (few row tiles), see pmb_qlearner_train_step
MixDwPlan mix_dw_plan(const pmb_dims* d, int ctas_avail = 0) {
    MixDwPlan p;
    p.n_ct = (tc_mix_cblks(d) + 3) / 4;
    p.n_kt = (tc_state_chunks(d) + 3) / 4;
    p.ldk = tc_state_chunks(d) * 64;
    p.n_halves = tc_mix_row_tiles(d) * 2;
    int64_t tiles = (int64_t)p.n_ct * p.n_kt;
    int64_t sl = (ctas_avail > 0 && ctas_avail < sm_count() ? ctas_avail : sm_count()) / tiles;
    if (sl > p.n_halves / 8) sl = p.n_halves / 8;
    if (sl < 1) sl = 1;
    p.slices = (int)sl;
    p.halves_per_slice = ceil_div(p.n_halves, p.slices);
    return p;
}
int mix_bwd_grid(const pmb_dims* d) {
    int64_t g = 6 * (int64_t)sm_count();          // two full waves at 3 resident CTAs per SM (8 x SMs was 2.67 waves)
    int64_t mx = ceil_div(tc_mix_row_tiles(d) * 128, tc::MB_WARPS);
    if (g > mx) g = mx;
    return (int)(g < 1 ? 1 : g);
}
int64_t mix_v2_partial_bytes(const pmb_dims* d) { return align_up((int64_t)mix_bwd_grid(d) * 33 * 4, 256); }
}  // namespace

int tc_mixer_dw_ctas(const pmb_dims* d) { return ((tc_mix_cblks(d) + 3) / 4) * ((tc_state_chunks(d) + 3) / 4); }

// the slice count only shrinks with ctas_avail: the full-device plan bounds every plan
int64_t tc_mixer_dw_scratch_bytes(const pmb_dims* d) {
    MixDwPlan p = mix_dw_plan(d);
    return align_up((int64_t)p.slices * p.n_ct * 256 * p.ldk * 4, 256);
}

int64_t tc_mixer_bwd_img_scratch_bytes(const pmb_dims* d) { return mix_v2_partial_bytes(d) + tc_mixer_dw_scratch_bytes(d); }

// part 1: d_raw in place, d_agent_qs, V.2 gradients (scratch: mix_v2_partial_bytes, the start of the step's scratch area)
int tc_mixer_bwd_img_dq(const pmb_dims* d, const MixerParams& mp, uint8_t* raw_img, const float* agent_qs, const float* g,
                        float* d_agent_qs, float* gv2_w, float* gv2_b, void* scratch, int64_t scratch_bytes, cudaStream_t s) {
    if (d->E != 32) { set_error("tc_mixer_bwd_img: mixing_embed_dim must be 32"); return PMB_ERR_INVALID; }
    if (scratch_bytes < mix_v2_partial_bytes(d)) { set_error("tc_mixer_bwd_img: scratch too small"); return PMB_ERR_WORKSPACE; }
    const int grid = mix_bwd_grid(d);
    float* v2_partial = static_cast<float*>(scratch);
    tc::MixBwdParams B;
    B.raw_img = raw_img; B.agent_qs = agent_qs; B.g = g; B.v2_w = mp.v2_w; B.d_qs = d_agent_qs; B.v2_partial = v2_partial;
    B.BT = (int64_t)d->B * d->T; B.rows_total = tc_mix_row_tiles(d) * 128; B.T = d->T; B.N = d->N; B.n_cblk = tc_mix_cblks(d);
    if (B.n_cblk <= 16 && B.rows_total < (1ll << 31)) {
        static const bool generic = []() { const char* e = getenv("PMB_MIXBWD_GENERIC"); return e && atoi(e) != 0; }();   // A/B switch
        switch (generic ? 0 : d->N) {
#define PMB_MB16(n) case n: tc::mix_bwd_img16_kernel<n><<<grid, 32 * tc::MB_WARPS, 0, s>>>(B); break;
            PMB_MB16(1) PMB_MB16(2) PMB_MB16(3) PMB_MB16(4) PMB_MB16(5) PMB_MB16(6) PMB_MB16(7) PMB_MB16(8) PMB_MB16(9)
            PMB_MB16(10) PMB_MB16(11) PMB_MB16(12) PMB_MB16(13) PMB_MB16(14) PMB_MB16(15) PMB_MB16(16) PMB_MB16(17)
            PMB_MB16(18) PMB_MB16(19) PMB_MB16(20) PMB_MB16(21) PMB_MB16(22) PMB_MB16(23) PMB_MB16(24) PMB_MB16(25)
            PMB_MB16(26) PMB_MB16(27) PMB_MB16(28) PMB_MB16(29)
#undef PMB_MB16
            default: tc::mix_bwd_img16_kernel<0><<<grid, 32 * tc::MB_WARPS, 0, s>>>(B); break;
        }
    } else tc::mix_bwd_img_kernel<34><<<grid, 32 * tc::MB_WARPS, 0, s>>>(B);
    PMB_LAUNCH_CHECK("mix_bwd_img_kernel");
    tc::mix_v2_reduce_kernel<<<33, 256, 0, s>>>(v2_partial, grid, gv2_w, gv2_b);
    PMB_LAUNCH_CHECK("mix_v2_reduce_kernel");
    return PMB_OK;
}

// part 2: hypernet weight / bias gradients = d_raw^T . [state | 1] (scratch: tc_mixer_dw_scratch_bytes)
int tc_mixer_dw(const pmb_dims* d, const uint8_t* state_img, const uint8_t* raw_img, float* gw_cat, float* gb_cat,
                void* scratch, int64_t scratch_bytes, int ctas_avail, cudaStream_t s) {
    if (scratch_bytes < tc_mixer_dw_scratch_bytes(d)) { set_error("tc_mixer_dw: scratch too small"); return PMB_ERR_WORKSPACE; }
    float* partial = static_cast<float*>(scratch);
    MixDwPlan p = mix_dw_plan(d, ctas_avail);
    tc::MixDwParams W;
    W.raw_img = raw_img; W.state_img = state_img; W.partial = partial;
    W.n_cblk = tc_mix_cblks(d); W.n_chunks = tc_state_chunks(d); W.n_ct = p.n_ct; W.n_kt = p.n_kt; W.ldk = p.ldk;
    W.n_halves = p.n_halves; W.halves_per_slice = p.halves_per_slice;
    PMB_SMEM_ATTR(tc::mix_dw_tc_kernel, tc::md::SMEM_BYTES);
    tc::mix_dw_tc_kernel<<<p.n_ct * p.n_kt * p.slices, tc::md::THREADS, tc::md::SMEM_BYTES, s>>>(W);
    PMB_LAUNCH_CHECK("mix_dw_tc_kernel");
    const int64_t total = (int64_t)(d->N + 3) * 32 * (d->S + 1);
    tc::mix_dw_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(partial, p.slices, p.n_ct * 256, p.ldk, d->N,
                                                                          d->S, gw_cat, gb_cat);
    PMB_LAUNCH_CHECK("mix_dw_reduce_kernel");
    return PMB_OK;
}

int tc_mixer_bwd_img(const pmb_dims* d, const MixerParams& mp, const uint8_t* state_img, uint8_t* raw_img,
                     const float* agent_qs, const float* g, float* d_agent_qs, float* gw_cat, float* gb_cat, float* gv2_w,
                     float* gv2_b, void* scratch, int64_t scratch_bytes, cudaStream_t s) {
    if (scratch_bytes < tc_mixer_bwd_img_scratch_bytes(d)) { set_error("tc_mixer_bwd_img: scratch too small"); return PMB_ERR_WORKSPACE; }
    int rc = tc_mixer_bwd_img_dq(d, mp, raw_img, agent_qs, g, d_agent_qs, gv2_w, gv2_b, scratch, scratch_bytes, s);
    if (rc) return rc;
    const int64_t off = mix_v2_partial_bytes(d);
    return tc_mixer_dw(d, state_img, raw_img, gw_cat, gb_cat, static_cast<char*>(scratch) + off, scratch_bytes - off, 0, s);
}

}  // namespace pmb
