#include <cstdlib>
#include <algorithm>
// bf16 tensor-core tier: C[m, n] = sum_k A[m, k] * W[n, k] on tcgen05 with fp32 accumulation in TMEM.
//
//   A  fp32 in global memory, addressed through RowMap (the EpisodeBatch fields are consumed in
//      place).  Eight producer warps stream it with coalesced loads, convert to bf16 in registers
//      and write the 128-byte-swizzled K-major operand tile into shared memory.
//   W  packed once per step (pack_w_kernel) into bf16 tiles that already are the shared-memory image
//      (K-major, 128B swizzle), so one elected thread brings a tile in with a single bulk async copy
//      (cp.async.bulk, completion on an mbarrier) - no tensor map needed.
//   D  128 x <=256 fp32 accumulator in TMEM, double buffered (2 x 256 columns): the four epilogue warps
//      drain pass p with tcgen05.ld while the MMA thread already accumulates pass p+1.
//
// Warp roles (448 threads): 0-3 epilogue (warp i owns TMEM lanes 32i..32i+31 = tile rows),
// 4 TMEM allocator + single-thread tcgen05.mma issuer, 5 W bulk-copy issuer, 6-13 A producers.
// Pipelines: 4 shared-memory stages (full/empty mbarriers), 2 accumulator buffers (tmem_full/tmem_empty).
// Persistent over 128-row tiles: grid = min(#tiles, #SMs).
#include "common.cuh"
#include "tc_common.cuh"
#include "mixer_tc.cuh"

namespace pmb {
namespace tc {

constexpr int BM = 128;              // rows per tile = UMMA M
constexpr int BK = 64;               // bf16 elements per k-chunk = one 128-byte swizzle row
constexpr int BN_MAX = 256;          // columns per pass = UMMA N
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * 128;          // 16 KB
constexpr int W_STAGE_BYTES = BN_MAX * 128;      // 32 KB
constexpr int N_EPI_WARPS = 4, MMA_WARP = 4, WLOAD_WARP = 5, FIRST_PROD_WARP = 6, N_PROD_WARPS = 8;
constexpr int THREADS = 32 * (FIRST_PROD_WARP + N_PROD_WARPS);     // 448
constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * (A_STAGE_BYTES + W_STAGE_BYTES) + 256;

struct GemmParams {
    const float* A;
    RowMap amap;
    int64_t M;
    int K;                 // real reduction length
    int n_chunks;          // ceil(K / 64)
    const __nv_bfloat16* Wp;   // packed tiles: pass-major, then k-chunk; each tile ncols_p rows x 128 B
    int Ncols;             // multiple of 32
    // time-major tiled row order (tm_R > 0): logical row m = t * tm_Rpad + p, p = b*N + n < tm_R is valid and
    // maps to batch row (b*tm_T + t)*tm_N + n of amap; rows p >= tm_R are padding.  A 128-row tile then is
    // exactly one (t, tile) tile image of the bf16 tier.
    int64_t tm_R, tm_Rpad;
    int tm_N, tm_T;
    uint8_t* a_img_out;    // optional: the converted A tiles are also written to global, [tile][k-chunk][16 KB]
    const uint8_t* a_img;  // optional: A is already bf16 tile images [tile][k-chunk][16 KB] -> bulk copied, no producers
};

__host__ __device__ inline int pass_cols(int Ncols, int p) {
    int left = Ncols - p * BN_MAX;
    return left < BN_MAX ? left : BN_MAX;
}
__host__ __device__ inline int64_t packed_tile_offset(int Ncols, int n_chunks, int p, int c) {   // in elements
    return ((int64_t)p * BN_MAX * n_chunks + (int64_t)c * pass_cols(Ncols, p)) * BK;
}

// ------------------------------------------------------------------------------------------
// W packing: up to 4 row segments of fp32 [rows, ld] matrices -> bf16 swizzled tiles
// ------------------------------------------------------------------------------------------
struct PackSegs {
    const float* ptr[4];
    int rows[4];
    int ld[4];
    int nseg;
};

__global__ void __launch_bounds__(256)
pack_w_kernel(PackSegs segs, int K, int n_chunks, int Ncols, __nv_bfloat16* __restrict__ out) {
    // one thread per 16-byte chunk (8 bf16) of the packed image
    const int64_t total = (int64_t)Ncols * n_chunks * 8;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int j = (int)(i & 7);
    int64_t rc = i >> 3;
    int col = (int)(rc % Ncols);
    int c = (int)(rc / Ncols);
    int p = col / BN_MAX, r = col - p * BN_MAX;
    const float* src = nullptr;
    int rr = col;
    for (int s = 0; s < segs.nseg; ++s) {
        if (rr < segs.rows[s]) { src = segs.ptr[s] ? segs.ptr[s] + (int64_t)rr * segs.ld[s] : nullptr; break; }
        rr -= segs.rows[s];
    }
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int k = c * BK + j * 8 + q * 2;
        float lo = (src && k < K) ? src[k] : 0.f;
        float hi = (src && k + 1 < K) ? src[k + 1] : 0.f;
        w[q] = pack_bf16x2(lo, hi);
    }
    char* tile = reinterpret_cast<char*>(out + packed_tile_offset(Ncols, n_chunks, p, c));
    *reinterpret_cast<uint4*>(tile + sw128_offset((uint32_t)r, (uint32_t)j)) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------
// the GEMM kernel
// ------------------------------------------------------------------------------------------
template <class Epi>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(GemmParams P, Epi epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_stage = smem;
    uint8_t* w_stage = smem + STAGES * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + W_STAGE_BYTES));
    uint64_t* full = bars;                 // [STAGES]  producers + W bytes landed
    uint64_t* empty = bars + STAGES;       // [STAGES]  MMAs that read the stage retired
    uint64_t* tfull = bars + 2 * STAGES;   // [2]       accumulator complete
    uint64_t* tempty = tfull + 2;          // [2]       accumulator drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], P.a_img ? 1 : N_PROD_WARPS + 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], N_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t n_tiles = (P.M + BM - 1) / BM;
    const int n_pass = (P.Ncols + BN_MAX - 1) / BN_MAX;

    if (warp >= FIRST_PROD_WARP) {
        // ===== A producers: fp32 global -> bf16 swizzled smem (idle when A comes as tile images) =====
        const int pw = warp - FIRST_PROD_WARP;
        uint32_t it = 0;
        if (P.a_img == nullptr)
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            // lane l (< 16) keeps the pointer of tile row (pw + 8 l)
            const float* myptr = nullptr;
            if (lane < BM / N_PROD_WARPS) {
                int64_t row = tile * BM + pw + N_PROD_WARPS * lane;
                if (row < P.M) {
                    if (P.tm_R > 0) {
                        const int64_t t = row / P.tm_Rpad, pp = row - t * P.tm_Rpad;
                        if (pp < P.tm_R) {
                            const int64_t bb = pp / P.tm_N, nn = pp - bb * P.tm_N;
                            myptr = P.A + P.amap.offset((bb * P.tm_T + t) * P.tm_N + nn);
                        }
                    } else {
                        myptr = P.A + P.amap.offset(row);
                    }
                }
            }
            for (int p = 0; p < n_pass; ++p) {
                for (int c = 0; c < P.n_chunks; ++c, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* dst = a_stage + s * A_STAGE_BYTES;
                    const int k = c * BK + 2 * lane;
                    float v0[BM / N_PROD_WARPS], v1[BM / N_PROD_WARPS];
#pragma unroll
                    for (int l = 0; l < BM / N_PROD_WARPS; ++l) {
                        const float* rp = reinterpret_cast<const float*>(
                            __shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(myptr), l));
                        float a = 0.f, b = 0.f;
                        if (rp != nullptr) {
                            if (k + 1 < P.K && ((reinterpret_cast<uintptr_t>(rp + k) & 7) == 0)) {
                                float2 t = __ldg(reinterpret_cast<const float2*>(rp + k));
                                a = t.x; b = t.y;
                            } else {
                                if (k < P.K) a = __ldg(rp + k);
                                if (k + 1 < P.K) b = __ldg(rp + k + 1);
                            }
                        }
                        v0[l] = a; v1[l] = b;
                    }
                    uint8_t* gimg = (P.a_img_out && p == 0)
                                        ? P.a_img_out + ((int64_t)tile * P.n_chunks + c) * A_STAGE_BYTES : nullptr;
#pragma unroll
                    for (int l = 0; l < BM / N_PROD_WARPS; ++l) {
                        const uint32_t r = pw + N_PROD_WARPS * l;
                        const uint32_t off = sw128_offset(r, lane >> 2) + (lane & 3) * 4;
                        const uint32_t w = pack_bf16x2(v0[l], v1[l]);
                        *reinterpret_cast<uint32_t*>(dst + off) = w;
                        if (gimg) *reinterpret_cast<uint32_t*>(gimg + off) = w;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                }
            }
        }
    } else if (warp == WLOAD_WARP) {
        // ===== W loader: one bulk copy per (pass, k-chunk) =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int p = 0; p < n_pass; ++p) {
                    const uint32_t bytes = (uint32_t)pass_cols(P.Ncols, p) * 128u;
                    for (int c = 0; c < P.n_chunks; ++c, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                        mbar_arrive_expect_tx(&full[s], bytes + (P.a_img ? A_STAGE_BYTES : 0));
                        bulk_copy_g2s(w_stage + s * W_STAGE_BYTES, P.Wp + packed_tile_offset(P.Ncols, P.n_chunks, p, c),
                                      bytes, &full[s]);
                        if (P.a_img)
                            bulk_copy_g2s(a_stage + s * A_STAGE_BYTES, P.a_img + ((int64_t)tile * P.n_chunks + c) * A_STAGE_BYTES,
                                          A_STAGE_BYTES, &full[s]);
                    }
                }
        }
    } else if (warp == MMA_WARP) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t it = 0, acc_it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int p = 0; p < n_pass; ++p, ++acc_it) {
                    const int b = acc_it & 1;
                    const uint32_t idesc = umma_idesc_bf16(BM, pass_cols(P.Ncols, p), 0, 0);
                    mbar_wait(&tempty[b], ((acc_it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)b * BN_MAX;
                    for (int c = 0; c < P.n_chunks; ++c, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(&full[s], (it / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(a_stage + s * A_STAGE_BYTES);
                        const uint32_t w_addr = smem_u32(w_stage + s * W_STAGE_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            umma_bf16(tmem_d, umma_desc_sw128(a_addr + kk * 32, 16, 1024),
                                      umma_desc_sw128(w_addr + kk * 32, 16, 1024), idesc, (c | kk) != 0);
                        }
                        umma_commit(&empty[s]);
                    }
                    umma_commit(&tfull[b]);
                }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> Epi =====
        uint32_t acc_it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t m = tile * BM + warp * 32 + lane;
            const bool valid = m < P.M;
            typename Epi::Row row;
            epi.begin(row, m, valid);
            epi.attach(row, smem + STAGES * (A_STAGE_BYTES + W_STAGE_BYTES) + 256 + warp * (Epi::EXTRA_SMEM / N_EPI_WARPS));
            for (int p = 0; p < n_pass; ++p, ++acc_it) {
                const int b = acc_it & 1;
                const int ncols = pass_cols(P.Ncols, p);
                mbar_wait(&tfull[b], (acc_it >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (uint32_t)b * BN_MAX + ((uint32_t)(warp * 32) << 16);
                for (int g = 0; g < ncols; g += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + g, v);
                    tmem_wait_ld();
                    epi.cols(row, m, valid, p * BN_MAX + g, v);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[b]);
            }
            epi.end(row, m, valid);
        }
        epi.finish();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// epilogues
// ------------------------------------------------------------------------------------------
struct PlainEpi {           // C[m*ldc + n] = acc + bias[n]
    static constexpr int EXTRA_SMEM = 0;
    template <class R> __device__ void attach(R&, uint8_t*) const {}
    __device__ void finish() const {}
    struct Row {};
    float* C; int64_t ldc; const float* bias; int Nreal;
    __device__ void begin(Row&, int64_t, bool) const {}
    __device__ void cols(Row&, int64_t m, bool valid, int col0, const uint32_t (&v)[32]) const {
        if (!valid) return;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j < Nreal) C[m * ldc + col0 + j] = __uint_as_float(v[j]) + (bias ? bias[col0 + j] : 0.f);
    }
    __device__ void end(Row&, int64_t, bool) const {}
};

// fc1 for the online and the target net in one GEMM: columns [0, 64) online, [64, 128) target.
//   x = relu(acc + tab_id[net][n][h] (bias folded in) + tab_act[net][a_prev][h])
struct Fc1Epi {
    static constexpr int EXTRA_SMEM = 0;
    template <class R> __device__ void attach(R&, uint8_t*) const {}
    __device__ void finish() const {}
    struct Row { int64_t out_off; int n; int a_prev; };
    const float* tab_act;       // [2][A][64]
    const float* tab_id;        // [2][N][64]   W_id[:, n] + b1 (or just b1 when obs_agent_id is off)
    const int64_t* actions; int64_t actions_sb;
    const int64_t* filled;  int64_t filled_sb;
    float* x_on; float* x_tg;   // [nt][R][64] time major
    int t0, nt, N, A, use_act;
    int64_t R;
    __device__ void begin(Row& r, int64_t m, bool valid) const {
        r.out_off = 0; r.n = 0; r.a_prev = -1;
        if (!valid) return;
        const int tn = nt * N;
        int64_t b = m / tn;
        int rem = (int)(m - b * tn);
        int tl = rem / N;
        r.n = rem - tl * N;
        int t = t0 + tl;
        if (use_act && t > 0 && filled[b * filled_sb + (t - 1)] != 0)
            r.a_prev = (int)actions[b * actions_sb + (int64_t)(t - 1) * N + r.n];
        r.out_off = ((int64_t)tl * R + b * N + r.n) * 64;
    }
    __device__ void cols(Row& r, int64_t, bool valid, int col0, const uint32_t (&v)[32]) const {
        if (!valid) return;
        const int net = col0 >> 6, h0 = col0 & 63;
        float* out = (net ? x_tg : x_on) + r.out_off + h0;
        const float4* tid = reinterpret_cast<const float4*>(tab_id + ((int64_t)net * N + r.n) * 64 + h0);
        const float4* tac = r.a_prev >= 0 ? reinterpret_cast<const float4*>(tab_act + ((int64_t)net * A + r.a_prev) * 64 + h0)
                                          : nullptr;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4 bi = __ldg(tid + q);
            float4 ac = tac ? __ldg(tac + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o;
            o.x = fmaxf(__uint_as_float(v[4 * q + 0]) + ac.x + bi.x, 0.f);
            o.y = fmaxf(__uint_as_float(v[4 * q + 1]) + ac.y + bi.y, 0.f);
            o.z = fmaxf(__uint_as_float(v[4 * q + 2]) + ac.z + bi.z, 0.f);
            o.w = fmaxf(__uint_as_float(v[4 * q + 3]) + ac.w + bi.w, 0.f);
            reinterpret_cast<float4*>(out)[q] = o;
        }
    }
    __device__ void end(Row&, int64_t, bool) const {}
};

// Same, but x is written as bf16 tile images [nt][n_tiles][16 KB] (the operand format of gru_tc.cu).
struct Fc1TiEpi {
    static constexpr int EXTRA_SMEM = 0;
    template <class R> __device__ void attach(R&, uint8_t*) const {}
    __device__ void finish() const {}
    struct Row { int64_t tile_off; uint32_t r; int n; int a_prev; };
    const float* tab_act; const float* tab_id;
    const int64_t* actions; int64_t actions_sb;
    const int64_t* filled;  int64_t filled_sb;
    uint8_t* x_on; uint8_t* x_tg;
    int t0, nt, N, A, use_act, n_tiles;
    int64_t R;
    uint32_t* relu_mask;        // optional [nt][n_tiles][2][128]: bit j of word (half, row) = online x[32*half + j] > 0
    // rows arrive in time-major tiled order: m = tl * (n_tiles*128) + p
    __device__ void begin(Row& r, int64_t m, bool valid) const {
        r.tile_off = 0; r.r = 0; r.n = 0; r.a_prev = -2;          // -2: padding row (nothing is written)
        if (!valid) return;
        const int64_t r_pad = (int64_t)n_tiles * 128;
        const int64_t tl = m / r_pad, p = m - tl * r_pad;
        if (p >= R) return;
        const int64_t b = p / N;
        r.n = (int)(p - b * N);
        r.a_prev = -1;
        const int t = t0 + (int)tl;
        if (use_act && t > 0 && filled[b * filled_sb + (t - 1)] != 0)
            r.a_prev = (int)actions[b * actions_sb + (int64_t)(t - 1) * N + r.n];
        r.tile_off = (tl * n_tiles + (p >> 7)) * 16384;
        r.r = (uint32_t)(p & 127);
    }
    __device__ void cols(Row& r, int64_t, bool valid, int col0, const uint32_t (&v)[32]) const {
        if (!valid || r.a_prev == -2) return;
        const int net = col0 >> 6, h0 = col0 & 63;
        uint8_t* tile = (net ? x_tg : x_on) + r.tile_off;
        const float4* tid = reinterpret_cast<const float4*>(tab_id + ((int64_t)net * N + r.n) * 64 + h0);
        const float4* tac = r.a_prev >= 0 ? reinterpret_cast<const float4*>(tab_act + ((int64_t)net * A + r.a_prev) * 64 + h0)
                                          : nullptr;
        uint32_t mbits = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {                 // 8 columns = one 16-byte chunk
            float o[8];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float4 bi = __ldg(tid + 2 * q + hh);
                float4 ac = tac ? __ldg(tac + 2 * q + hh) : make_float4(0.f, 0.f, 0.f, 0.f);
                o[4 * hh + 0] = fmaxf(__uint_as_float(v[8 * q + 4 * hh + 0]) + ac.x + bi.x, 0.f);
                o[4 * hh + 1] = fmaxf(__uint_as_float(v[8 * q + 4 * hh + 1]) + ac.y + bi.y, 0.f);
                o[4 * hh + 2] = fmaxf(__uint_as_float(v[8 * q + 4 * hh + 2]) + ac.z + bi.z, 0.f);
                o[4 * hh + 3] = fmaxf(__uint_as_float(v[8 * q + 4 * hh + 3]) + ac.w + bi.w, 0.f);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) mbits |= (o[k] > 0.f ? 1u : 0u) << (8 * q + k);
            *reinterpret_cast<uint4*>(tile + sw128_offset(r.r, (uint32_t)(h0 / 8 + q))) =
                make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        }
        if (relu_mask && net == 0) relu_mask[((r.tile_off >> 14) * 2 + (h0 >> 5)) * 128 + r.r] = mbits;
    }
    __device__ void end(Row&, int64_t, bool) const {}
};

// QMIX mixing on the hypernet outputs, E = 32.  Packed column order: [ w1 (N*32) | b1 | w_final | v0 ].
//   hidden = ELU(sum_n q[n] |w1[n, :]| + b1);  q_tot = hidden . |w_final| + V.2(ReLU(v0))
// Mixer epilogue for the image-fed path: GEMM rows are ALL (b, t) pairs, m' = b*T + t.  The online mixer
// (t_off = 0) uses rows t < T-1, the target mixer (t_off = 1) rows t >= 1; agent_qs / q_tot are indexed by
// mq = b*(T-1) + t - t_off.  raw_img (online only): hypernet outputs as bf16 tile images
// [row tile][(N+3)/2 column blocks of 64][16 KB], packed column order, kept for the backward.
struct MixImgEpi {
    // raw images leave through shared memory: every epilogue warp stages its 32 rows of a 64-column block (4 KB, image
    // layout) in its own double buffer and lane 0 bulk-stores it - no per-thread row-strided global stores
    static constexpr int EXTRA_SMEM = N_EPI_WARPS * 2 * 4096;
    struct Row { float hidden[32]; float y; int64_t mq; uint8_t* img; uint32_t r; uint8_t* stg; uint32_t nblk; };
    const float* bias; const float* agent_qs; const float* v2_w; const float* v2_b;
    float* q_tot; uint8_t* raw_img;
    int N, T, t_off, n_cblk;
    int64_t BT;
    __device__ void begin(Row& r, int64_t m, bool) const {
#pragma unroll
        for (int e = 0; e < 32; ++e) r.hidden[e] = 0.f;
        r.y = 0.f; r.mq = -1;
        // this warp's 32 rows of the tile: 4 KB inside every 16 KB block
        r.img = raw_img ? raw_img + (m >> 7) * (int64_t)n_cblk * 16384 + ((m & 127) & ~31) * 128 : nullptr;
        r.r = (uint32_t)(m & 127);
        r.nblk = 0;
        if (m < BT) {
            const int64_t b = m / T;
            const int t = (int)(m - b * T);
            if (t_off == 0 ? (t < T - 1) : (t >= 1)) r.mq = b * (T - 1) + t - t_off;
        }
    }
    __device__ void attach(Row& r, uint8_t* stg) const { r.stg = stg; }
    __device__ void flush(Row& r, int blk) const {              // whole warp
        fence_proxy_async_smem();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            bulk_copy_s2g(r.img + (int64_t)blk * 16384, r.stg + (r.nblk & 1) * 4096, 4096);
            bulk_commit_group();
            bulk_wait_group_read<1>();                          // the other buffer (previous block) has been read
        }
        __syncwarp();
        ++r.nblk;
    }
    __device__ void cols(Row& r, int64_t, bool, int col0, const uint32_t (&v)[32]) const {
        const int grp = col0 >> 5;
        float raw[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) raw[e] = __uint_as_float(v[e]) + __ldg(bias + col0 + e);
        if (r.img) {                                   // every row is written (finite values; the backward zeroes unused rows)
            uint8_t* blk = r.stg + (r.nblk & 1) * 4096;
            const uint32_t ch0 = (uint32_t)((col0 & 63) >> 3);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(blk + sw128_offset(r.r & 31u, ch0 + q)) =
                    make_uint4(pack_bf16x2(raw[8 * q], raw[8 * q + 1]), pack_bf16x2(raw[8 * q + 2], raw[8 * q + 3]),
                               pack_bf16x2(raw[8 * q + 4], raw[8 * q + 5]), pack_bf16x2(raw[8 * q + 6], raw[8 * q + 7]));
            if (col0 & 32) flush(r, col0 >> 6);        // second half of a 64-column block
        }
        if (grp < N) {
            const float qn = r.mq >= 0 ? __ldg(agent_qs + r.mq * N + grp) : 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) r.hidden[e] = fmaf(qn, fabsf(raw[e]), r.hidden[e]);
        } else if (grp == N) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                float pre = r.hidden[e] + raw[e];
                r.hidden[e] = pre > 0.f ? pre : expm1f(pre);
            }
        } else if (grp == N + 1) {
#pragma unroll
            for (int e = 0; e < 32; ++e) r.y = fmaf(r.hidden[e], fabsf(raw[e]), r.y);
        } else if (grp == N + 2) {
#pragma unroll
            for (int e = 0; e < 32; ++e) r.y = fmaf(fmaxf(raw[e], 0.f), __ldg(v2_w + e), r.y);
        }
    }
    __device__ void end(Row& r, int64_t, bool) const {
        if (r.mq >= 0) q_tot[r.mq] = r.y + __ldg(v2_b);
        // (N + 3) odd: the last 64-column block was only half produced (its other half is padding the backward ignores)
        if (r.img) {
            if ((N + 3) & 1) flush(r, (N + 3) >> 1);
            // the next tile starts again with buffer 0: both buffers must have been read
            if ((threadIdx.x & 31) == 0) bulk_wait_group_read<0>();
            __syncwarp();
        }
    }
    __device__ void finish() const {
        if ((threadIdx.x & 31) == 0) bulk_wait_group<0>();
    }
};

template <class Epi>
int launch_tc_gemm(const GemmParams& P, const Epi& epi, cudaStream_t s) {
    if (P.M <= 0) return PMB_OK;
    auto kern = tc_gemm_kernel<Epi>;
    PMB_SMEM_ATTR(kern, SMEM_BYTES + Epi::EXTRA_SMEM);
    int64_t n_tiles = (P.M + BM - 1) / BM;
    int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    kern<<<grid, THREADS, SMEM_BYTES + Epi::EXTRA_SMEM, s>>>(P, epi);
    PMB_LAUNCH_CHECK("tc_gemm_kernel");
    return PMB_OK;
}


// ------------------------------------------------------------------------------------------
// fc1 for both nets as a STREAMING kernel (obs is 70 % of the learner step's HBM bytes)
// ------------------------------------------------------------------------------------------
// The generic kernel above feeds A through eight producer warps that load 8 bytes per lane (obs rows are 4-byte
// aligned only: 285 floats) and convert synchronously, chunk by chunk: ~28 % of the HBM ceiling (ncu, profiles/).
// Here the copy engine does the streaming:
//   producer warp : walks the rows of a 16-row group (time-major tile order: a group is 1-2 contiguous runs of
//                   [B,T,N,O] memory, one per episode), issues ONE cp.async.bulk per run for its 16-byte-aligned
//                   interior and 4-byte cp.async for the <= 3 floats of head and tail, into a ring of fp32 staging
//                   slots (never reads outside the run).  3 slots x 18 KB in flight per SM.
//   8 converter warps : fp32 staging -> bf16 K-major/128B-swizzle A tile (all k-chunks of the tile, 80 KB).
//   MMA thread    : W (both nets, 128 columns, all chunks) stays resident; 4 tcgen05.mma per chunk, 2 TMEM buffers.
//   store thread  : the finished A tile IS the obs tile image the weight-gradient kernel wants: one 80 KB bulk store.
//   4 epilogue warps : Fc1TiEpi (one-hot gathers, bias, ReLU, bf16 x tile images of both nets, ReLU bit mask).
namespace fs {
constexpr int MAX_CHUNKS = 5;
constexpr int GROUP_ROWS = 16, GROUPS = BM / GROUP_ROWS;     // (8 rows x 6 slots measured no better: 8.5 vs 8.2 ms)
constexpr int N_SLOTS = 3;                                   // learner step (both nets: W is 16 KB per k-chunk)
constexpr int N_SLOTS_1NET = 5;                              // rollout step: W of one net is half the size, two more slots fit
#ifndef PMB_FC1_NCONV
#define PMB_FC1_NCONV 8
#endif
constexpr int N_EPI = 4, MMA_W = 4, PROD_W = 5, N_CONV = PMB_FC1_NCONV;
constexpr int RPW = GROUP_ROWS / N_CONV;                     // rows of a group per converter warp
// NS staging slots, one producer warp each (every producer warp owns one staging slot: parity waits are only safe one
// phase apart), then the converter warps
constexpr int threads_for(int ns) { return 32 * (PROD_W + ns + N_CONV); }
static_assert(RPW >= 1 && RPW * N_CONV == GROUP_ROWS, "group rows must divide over the converter warps");
}  // namespace fs

struct Fc1StreamParams {
    const int64_t* ep_index;       // optional batch row -> buffer episode
    const float* obs; int64_t obs_sb;
    const __nv_bfloat16* Wp;       // packed [chunk][128 x 128 B]
    uint8_t* obs_img;              // optional [T*n_tiles][n_chunks][16 KB]
    int64_t R;
    int T, N, O, n_tiles, n_chunks, slot_bytes;
    // fold_id: the K padding of the last chunk carries one-hot(agent id) in columns O .. O+N-1 and 1.0 in column O+N;
    // the packed W holds the agent-id columns of fc1.weight and fc1.bias there, so the tensor core adds them.
    int fold_id;
    // epilogue
    const uint4* tab_act16;        // bf16 [A][net 2][64]: fc1.weight[:, O + a] of both nets
    const float* tab_id;           // fp32 [net 2][N][64] (used when !fold_id): W_id[:, n] + b1
    const int64_t* actions; int64_t actions_sb;
    const int64_t* filled;  int64_t filled_sb;
    uint8_t* x_on; uint8_t* x_tg;  // tile images [T][n_tiles][16 KB]
    uint32_t* relu_mask;           // [T][n_tiles][2][128]
    int use_act;
    int l2_prefetch;               // n > 0: prefetch the runs of the CTA's n-th next tile into L2 (off by default)
    int n_nets;                    // 2: online + target (learner step), 1: online only (rollout step)
    int t0;                        // batch timestep of item t = 0 (rollout: the step index; obs is addressed at t0 + t)
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// NC = number of 64-column k-chunks of the obs (compile time: the converters are instruction-latency bound - 8 warps,
// ~300 dependent instructions per 16-row group each - and with NC known the per-chunk predicates of all but the last
// chunk and the run-time "is this the last chunk" selects disappear)
// Per-role wait accounting (build with PMB_EXTRA_NVCC_FLAGS=-DPMB_FC1_PROFILE, read with tools/fc1_role_profile.py): cycles
// each role spends blocked on each barrier / in each converter phase, summed over CTAs.  This is what located the
// obs-image bulk store as the reason the converters idled between tiles.
#ifdef PMB_FC1_PROFILE
__device__ unsigned long long g_fc1_prof[16];
#define TWAIT(idx, bar, par) do { long long _t0 = clock64(); mbar_wait(bar, par); if (lane == 0) prof_acc[idx] += clock64() - _t0; } while (0)
#define TMARK(var) const long long var = clock64()
#define TADD(idx, expr) do { if (lane == 0) prof_acc[idx] += (expr); } while (0)
#else
#define TWAIT(idx, bar, par) mbar_wait(bar, par)
#define TMARK(var)
#define TADD(idx, expr)
#endif
template <int NC, int NS>
__global__ void __launch_bounds__(fs::threads_for(NS), 1) fc1_stream_kernel(Fc1StreamParams P) {
#ifdef PMB_FC1_PROFILE
    long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long prof_t0 = clock64();
#endif
    using namespace fs;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int N_SLOTS = NS, N_PROD = NS, FIRST_CONV_W = PROD_W + NS;
    constexpr int tile_bytes = NC * A_STAGE_BYTES;
    // W of a k-chunk: 128 rows x 128 B (online | target); the rollout step keeps the online half only
    const int w_chunk = P.n_nets == 2 ? A_STAGE_BYTES : A_STAGE_BYTES / 2;
    uint8_t* w_s = smem;                                    // [n_chunks][w_chunk]
    uint8_t* a_s = smem + NC * w_chunk;                     // [n_chunks][16 KB]
    uint8_t* stage = a_s + tile_bytes;                      // [N_SLOTS][slot_bytes]
    int* rowoff = reinterpret_cast<int*>(stage + N_SLOTS * P.slot_bytes);   // [N_SLOTS][16] byte offset of a row in its slot
    int* rown = rowoff + N_SLOTS * GROUP_ROWS;                              // [N_SLOTS][16] agent index of the row
    uint64_t* bars = reinterpret_cast<uint64_t*>(rown + N_SLOTS * GROUP_ROWS);
    uint64_t* w_full = bars;
    uint64_t* st_full = bars + 1;            // [N_SLOTS]
    uint64_t* st_empty = st_full + N_SLOTS;  // [N_SLOTS]
    uint64_t* a_full = st_empty + N_SLOTS;
    uint64_t* a_free = a_full + 1;
    uint64_t* tfull = a_free + 1;            // [2]
    uint64_t* tempty = tfull + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        for (int i = 0; i < N_SLOTS; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], N_CONV); }
        mbar_init(a_full, N_CONV); mbar_init(a_free, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], N_EPI); }
        fence_barrier_init();
    }
    if (warp == MMA_W) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t n_items = (int64_t)P.T * P.n_tiles;
    // programmatic dependent launch (rollout step only; see gru_rollout_kernel): the next kernel may be scheduled as the
    // CTAs of this grid leave; this kernel reads nothing its predecessor writes and only its epilogue warps write something
    // the predecessor (the GRU step of the previous rollout step) may still be reading - the x images
    pdl_launch_dependents();

    if (warp >= PROD_W && warp < PROD_W + N_PROD) {
        // one producer warp per staging slot: the address walk of a warp is ~300 dependent integer instructions per
        // group (~1.5 us); ONE warp capped the stream at 2.9 TB/s, two at 4.6 TB/s.  A slot is always filled by the
        // same warp, so consecutive uses of a slot are program-ordered (a parity wait must never be two phases ahead).
        const int pw = warp - PROD_W;
        if (pw == 0 && lane == 0) {
            mbar_arrive_expect_tx(w_full, (uint32_t)(NC * w_chunk));
            for (int c = 0; c < NC; ++c)
                bulk_copy_g2s(w_s + c * w_chunk, reinterpret_cast<const uint8_t*>(P.Wp) + c * A_STAGE_BYTES, (uint32_t)w_chunk, w_full);
        }
        // Lane i owns the i-th contiguous run of a group (one episode each) and computes its descriptor - global
        // address, staging offset, head / interior / tail split - in closed form, in parallel with the other lanes
        // and BEFORE the slot is free; once it is, every lane just issues its copies.  (The first version walked the
        // runs serially after the wait: ~1.5 us of dependent integer code per group during which the slot sat empty,
        // which bounded the stream at three slots per (walk + load + convert).)
        uint32_t git = 0;
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int64_t t = item / P.n_tiles, tile = item - t * P.n_tiles;
            for (int g = 0; g < GROUPS; ++g, ++git) {
                const int slot = git % N_SLOTS;
                if (slot != pw) continue;
                TMARK(pp0);
                uint8_t* sl = stage + slot * P.slot_bytes;
                int* ro = rowoff + slot * GROUP_ROWS;
                int* rn = rown + slot * GROUP_ROWS;
                const int64_t p0 = tile * BM + g * GROUP_ROWS;
                int rows_here = P.R - p0 < GROUP_ROWS ? (int)(P.R - p0) : GROUP_ROWS;     // rows of this group below R
                if (rows_here < 0) rows_here = 0;
                const uint32_t b0 = (uint32_t)p0 / (uint32_t)P.N;               // B*N < 2^31 (validated on the host)
                const int n00 = (int)((uint32_t)p0 - b0 * (uint32_t)P.N);
                int first_cnt = P.N - n00;
                if (first_cnt > rows_here) first_cnt = rows_here;
                // run `lane`: rows [lr, lr + cnt) of the group, agent index n0 .. of episode b0 + lane
                const int lr = lane == 0 ? 0 : first_cnt + (lane - 1) * P.N;
                int cnt = lane == 0 ? first_cnt : (rows_here - lr < P.N ? rows_here - lr : P.N);
                if (cnt < 0) cnt = 0;
                const int n0 = lane == 0 ? n00 : 0;
                const char* ga = nullptr;
                uint32_t cur = 0, phase = 0, cp_bytes = 0;
                if (cnt > 0) {
                    ga = reinterpret_cast<const char*>(P.obs + ep_row(P.ep_index, (int64_t)b0 + lane) * P.obs_sb +
                                                       ((t + P.t0) * P.N + n0) * (int64_t)P.O);
                    phase = (uint32_t)(reinterpret_cast<uintptr_t>(ga) & 15);
                    // staging offset: same 128-BYTE phase as the source (the copy engine then moves whole lines: with only the
                    // 16-byte phase matched every line is split and shifted); 320 bytes of slack per run keep the runs - and
                    // the up to 15 bytes copied before / after each of them - disjoint
                    cur = (((uint32_t)lr * P.O * 4u + 127u) & ~127u) + 320u * lane + (uint32_t)(reinterpret_cast<uintptr_t>(ga) & 127);
                    // ONE bulk copy per run, widened to 16-byte boundaries on both sides (the extra bytes stay inside the
                    // 16-byte granules the run touches anyway - never another page - and land in the slack).  The first
                    // version copied the aligned interior and fetched head / tail with 4-byte cp.async: with the 32
                    // no-increment barrier arrivals that needs, publishing a slot took ~1600 cycles.
                    cp_bytes = (phase + (uint32_t)cnt * P.O * 4u + 15u) & ~15u;
                }
                const uint32_t tx = __reduce_add_sync(0xffffffffu, cp_bytes);
                // L2 prefetch of the same group of this CTA's NEXT tile: the staging ring is all the shared memory that is
                // left (3 x 18 KB in flight per SM), so the copies should see L2 latency, not loaded-HBM latency
                if (P.l2_prefetch && cnt > 0 && item + (int64_t)P.l2_prefetch * gridDim.x < n_items) {
                    const int64_t item2 = item + (int64_t)P.l2_prefetch * gridDim.x;
                    const int64_t t2 = item2 / P.n_tiles, tile2 = item2 - t2 * P.n_tiles;
                    const int64_t q0 = tile2 * BM + g * GROUP_ROWS;
                    int rows2 = P.R - q0 < GROUP_ROWS ? (int)(P.R - q0) : GROUP_ROWS;
                    if (rows2 < 0) rows2 = 0;
                    const uint32_t c0 = (uint32_t)q0 / (uint32_t)P.N;
                    const int m00 = (int)((uint32_t)q0 - c0 * (uint32_t)P.N);
                    int fc = P.N - m00;
                    if (fc > rows2) fc = rows2;
                    const int lr2 = lane == 0 ? 0 : fc + (lane - 1) * P.N;
                    int cnt2 = lane == 0 ? fc : (rows2 - lr2 < P.N ? rows2 - lr2 : P.N);
                    if (cnt2 > 0) {
                        const char* gb = reinterpret_cast<const char*>(P.obs + ep_row(P.ep_index, (int64_t)c0 + lane) * P.obs_sb +
                                                                       ((t2 + P.t0) * P.N + (lane == 0 ? m00 : 0)) * (int64_t)P.O);
                        const uintptr_t a0 = (reinterpret_cast<uintptr_t>(gb) + 15) & ~(uintptr_t)15;
                        const uintptr_t a1 = (reinterpret_cast<uintptr_t>(gb) + (uint64_t)cnt2 * P.O * 4) & ~(uintptr_t)15;
                        if (a1 > a0)
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((uint32_t)(a1 - a0)) : "memory");
                    }
                }
                // row tables, one row per lane: row j belongs to run rj (run 0 holds first_cnt rows, the others N each)
                const int rj = lane < first_cnt ? 0 : 1 + (lane - first_cnt) / P.N;
                const int lrj = rj == 0 ? 0 : first_cnt + (rj - 1) * P.N;
                const uint32_t cur_j = __shfl_sync(0xffffffffu, cur, rj & 31);
                TADD(10, clock64() - pp0);                         // producer: descriptors of the group (before the slot wait)
                TWAIT(0, &st_empty[slot], ((git / N_SLOTS) & 1) ^ 1);
                TMARK(pp1);
                if (lane < GROUP_ROWS) {
                    ro[lane] = lane < rows_here ? (int)(cur_j + (uint32_t)(lane - lrj) * P.O * 4u) : -1;   // rows beyond R: zeros
                    rn[lane] = (rj == 0 ? n00 : 0) + (lane - lrj);
                }
                if (cnt > 0) bulk_copy_g2s(sl + cur - phase, ga - phase, cp_bytes, &st_full[slot]);
                __syncwarp();
                if (lane == 0) {
                    if (tx) mbar_arrive_expect_tx(&st_full[slot], tx); else mbar_arrive(&st_full[slot]);
                }
                TADD(11, clock64() - pp1);                         // producer: tables, copies, arrival
            }
        }
    } else if (warp >= FIRST_CONV_W) {
        // ===== converters: warp cw owns rows 2cw, 2cw+1 of every group.  Everything that does not depend on the row
        // (which of a lane's two columns per chunk exist, where the one-hot padding columns are) is hoisted: the
        // first version spent ~2400 cycles per group in a serial chain of predicate / address arithmetic. =====
        const int cw = warp - FIRST_CONV_W;
        constexpr int c_last = NC - 1;
        // every column of the chunks before the last exists (O > 64 (NC-1)); the last chunk: per lane
        const bool okl0 = c_last * BK + 2 * lane < P.O, okl1 = c_last * BK + 2 * lane + 1 < P.O;
        const int j_pad = c_last * BK + 2 * lane - P.O;       // index of this lane's first column in the K padding
        const uint32_t a_s_u32 = smem_u32(a_s);
        const uint32_t a_base = a_s_u32 + ((uint32_t)(lane >> 2) << 4) + (lane & 3) * 4;   // + row*128, ^ swizzle
        uint32_t git = 0, ti = 0;
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++ti) {
            for (int g = 0; g < GROUPS; ++g, ++git) {
                const int slot = git % N_SLOTS;
                TWAIT(1, &st_full[slot], (git / N_SLOTS) & 1);
                TMARK(pt1);
                const uint32_t sl = smem_u32(stage + slot * P.slot_bytes) + 8u * lane;
                int offs[RPW], nns[RPW];
#pragma unroll
                for (int rr = 0; rr < RPW; ++rr) {
                    offs[rr] = rowoff[slot * GROUP_ROWS + RPW * cw + rr];
                    nns[rr] = rown[slot * GROUP_ROWS + RPW * cw + rr];
                }
                float v0[RPW][NC], v1[RPW][NC];
                {
#pragma unroll
                    for (int rr = 0; rr < RPW; ++rr) {           // all loads first
                        const int off = offs[rr];
                        const bool real = off >= 0;
                        const uint32_t src = sl + (uint32_t)(real ? off : 0);
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            v0[rr][c] = 0.f; v1[rr][c] = 0.f;
                            if (real && (c < c_last || okl0)) v0[rr][c] = lds_f32(src + c * 256);
                            if (real && (c < c_last || okl1)) v1[rr][c] = lds_f32(src + c * 256 + 4);
                        }
                    }
                    // pack first (consumes every loaded value), release the staging slot, then write the A tile
                    uint32_t pk[RPW][NC];
#pragma unroll
                    for (int rr = 0; rr < RPW; ++rr) {
                        const int off = offs[rr];
                        const int nn = nns[rr];
                        // one-hot(agent) and the bias column live in the K padding of the last chunk
                        const bool fold = P.fold_id && off >= 0;
                        const bool hit0 = fold && j_pad >= 0 && (j_pad == nn || j_pad == P.N);
                        const bool hit1 = fold && j_pad + 1 >= 0 && (j_pad + 1 == nn || j_pad + 1 == P.N);
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            float a = v0[rr][c], b = v1[rr][c];
                            if (c == c_last) { a = hit0 ? 1.0f : a; b = hit1 ? 1.0f : b; }
                            pk[rr][c] = pack_bf16x2(a, b);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&st_empty[slot]);
                    TMARK(pt2);
                    TADD(6, pt2 - pt1);                            // load + pack phase
                    // the first group of a tile waits here (values already in registers, slot already released) until
                    // the MMAs and the image store of the previous tile are done with the A tile
                    if (g == 0) TWAIT(2, a_free, (ti & 1) ^ 1);
#pragma unroll
                    for (int rr = 0; rr < RPW; ++rr) {
                        const uint32_t rt = (uint32_t)(g * GROUP_ROWS + RPW * cw + rr);
                        const uint32_t dst = (a_base + rt * 128u) ^ ((rt & 7u) << 4);
#pragma unroll
                        for (int c = 0; c < NC; ++c) sts_b32(dst + c * A_STAGE_BYTES, pk[rr][c]);
                        // the obs image leaves from here too: the warp's 32 words are one full 128-byte line of the image
                        // (a bulk store of the finished A tile kept the tile busy for ~2500 cycles after the MMAs were
                        // done with it: the converters of the next tile waited 22 % of the kernel time for it)
                        if (P.obs_img) {
                            uint8_t* img = P.obs_img + item * tile_bytes + (dst - a_s_u32);
#pragma unroll
                            for (int c = 0; c < NC; ++c) __stcs(reinterpret_cast<uint32_t*>(img + c * A_STAGE_BYTES), pk[rr][c]);
                        }
                    }
                    TADD(7, clock64() - pt2);                      // store phase (includes the a_free wait of group 0)
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
    } else if (warp == MMA_W) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(BM, 64 * P.n_nets, 0, 0);
            mbar_wait(w_full, 0);
            uint32_t ti = 0;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++ti) {
                const int b = ti & 1;
                TWAIT(3, a_full, ti & 1);
                TWAIT(4, &tempty[b], ((ti >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tm = tmem_base + 128u * b;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const uint32_t a_addr = smem_u32(a_s + c * A_STAGE_BYTES), w_addr = smem_u32(w_s + c * w_chunk);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk)
                        umma_bf16(tm, umma_desc_sw128(a_addr + kk * 32, 16, 1024), umma_desc_sw128(w_addr + kk * 32, 16, 1024),
                                  idesc, (c | kk) != 0);
                }
                umma_commit(a_free);
                umma_commit(&tfull[b]);
            }
        }
    } else {
        // (no store warp: the obs image is written by the converters.  One warp copying the finished A tile out with 16-byte
        // loads / stores took longer than the MMAs - 8.65 against 7.23 ms)
        // ===== epilogue: one row per thread; x = relu(acc [+ id/bias table] + act table) -> bf16 tile images + mask.
        // (Eight epilogue warps, one per row and net, measured slower: 7.21 -> 7.53 ms.) =====
        const uint32_t r = (uint32_t)(warp * 32 + lane);
        // the previous action of a row, fetched one item ahead (-1: none / step not filled, -2: padding row)
        auto fetch_prev = [&](int64_t item, int& n_out) {
            n_out = 0;
            if (item >= n_items) return -2;
            const int64_t t = item / P.n_tiles, tile = item - t * P.n_tiles;
            const int64_t p = tile * BM + r;
            if (p >= P.R) return -2;
            const int64_t b = p / P.N;
            n_out = (int)(p - b * P.N);
            const int64_t tb = t + P.t0;                       // timestep in the batch
            if (!P.use_act || tb == 0) return -1;
            const int64_t be = ep_row(P.ep_index, b);
            const int64_t f = __ldg(P.filled + be * P.filled_sb + (tb - 1));
            const int a = (int)__ldg(P.actions + be * P.actions_sb + (tb - 1) * P.N + n_out);
            return f != 0 ? a : -1;
        };
        int n_cur, n_nxt;
        int a_prev = fetch_prev(blockIdx.x, n_cur);
        uint32_t ti = 0;
        pdl_wait();                                            // before the first store to the x images
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++ti) {
            const int b = ti & 1;
            const int a_next = fetch_prev(item + gridDim.x, n_nxt);
            TWAIT(5, &tfull[b], (ti >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + 128u * b + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int g = 0; g < 4; ++g) {                      // 32 accumulator columns: net = g / 2, h0 = 32 (g & 1)
                uint32_t v[32];
                tmem_ld_32x32(taddr + 32 * g, v);
                tmem_wait_ld();
                if (a_prev != -2 && g < 2 * P.n_nets) {
                    const int net = g >> 1, h0 = (g & 1) * 32;
                    uint8_t* tile_img = (net ? P.x_tg : P.x_on) + item * A_STAGE_BYTES;
                    const uint4* tac = a_prev >= 0 ? P.tab_act16 + ((int64_t)a_prev * 2 + net) * 8 + (h0 >> 3) : nullptr;
                    const float4* tid = P.fold_id ? nullptr
                                                  : reinterpret_cast<const float4*>(P.tab_id + ((int64_t)net * P.N + n_cur) * 64 + h0);
                    uint32_t mbits = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {              // 8 columns = one 16-byte chunk
                        float o[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) o[k] = __uint_as_float(v[8 * q + k]);
                        if (tac) {
                            const uint4 w = __ldg(tac + q);
                            o[0] += __uint_as_float(w.x << 16); o[1] += __uint_as_float(w.x & 0xffff0000u);
                            o[2] += __uint_as_float(w.y << 16); o[3] += __uint_as_float(w.y & 0xffff0000u);
                            o[4] += __uint_as_float(w.z << 16); o[5] += __uint_as_float(w.z & 0xffff0000u);
                            o[6] += __uint_as_float(w.w << 16); o[7] += __uint_as_float(w.w & 0xffff0000u);
                        }
                        if (tid) {
                            const float4 b0 = __ldg(tid + 2 * q), b1 = __ldg(tid + 2 * q + 1);
                            o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
                            o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            o[k] = fmaxf(o[k], 0.f);
                            mbits |= (o[k] > 0.f ? 1u : 0u) << (8 * q + k);
                        }
                        *reinterpret_cast<uint4*>(tile_img + sw128_offset(r, (uint32_t)(h0 / 8 + q))) =
                            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                       pack_bf16x2(o[6], o[7]));
                    }
                    if (P.relu_mask && net == 0) P.relu_mask[(item * 2 + (h0 >> 5)) * 128 + r] = mbits;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);
            a_prev = a_next; n_cur = n_nxt;
        }
    }
#ifdef PMB_FC1_PROFILE
    if (lane == 0) {
        for (int i = 0; i < 8; ++i) if (prof_acc[i]) atomicAdd(&g_fc1_prof[i], (unsigned long long)prof_acc[i]);
        for (int i = 10; i < 12; ++i) if (prof_acc[i]) atomicAdd(&g_fc1_prof[i], (unsigned long long)prof_acc[i]);
        if (threadIdx.x == 0) { atomicAdd(&g_fc1_prof[8], (unsigned long long)(clock64() - prof_t0)); atomicAdd(&g_fc1_prof[9], 1ull); }
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_W) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------
// host entry points used by api.cu
// ------------------------------------------------------------------------------------------
int64_t tc_packed_elems(int Ncols, int K) {
    int n_chunks = (K + tc::BK - 1) / tc::BK;
    return (int64_t)Ncols * n_chunks * tc::BK;
}

int tc_pack_w(const float* const* ptrs, const int* rows, const int* lds, int nseg, int K, __nv_bfloat16* out,
              cudaStream_t s, int n_chunks_min) {
    tc::PackSegs segs;
    int Ncols = 0;
    segs.nseg = nseg;
    for (int i = 0; i < 4; ++i) {
        segs.ptr[i] = i < nseg ? ptrs[i] : nullptr;
        segs.rows[i] = i < nseg ? rows[i] : 0;
        segs.ld[i] = i < nseg ? lds[i] : 0;
        Ncols += segs.rows[i];
    }
    Ncols = (int)align_up(Ncols, 32);
    int n_chunks = (K + tc::BK - 1) / tc::BK;
    if (n_chunks < n_chunks_min) n_chunks = n_chunks_min;      // extra all-zero k-chunks (A images may be wider than K)
    int64_t total = (int64_t)Ncols * n_chunks * 8;
    tc::pack_w_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(segs, K, n_chunks, Ncols, out);
    PMB_LAUNCH_CHECK("pack_w_kernel");
    return PMB_OK;
}

// plain C = A . W^T + bias  (diagnostics / unit test of the tcgen05 pipeline)
int tc_gemm_plain(const float* A, RowMap amap, int64_t M, int K, const __nv_bfloat16* Wp, int Ncols_padded, int Nreal,
                  const float* bias, float* C, int64_t ldc, cudaStream_t s) {
    tc::GemmParams P{A, amap, M, K, (K + tc::BK - 1) / tc::BK, Wp, Ncols_padded, 0, 0, 0, 0, nullptr, nullptr};
    tc::PlainEpi epi{C, ldc, bias, Nreal};
    return tc::launch_tc_gemm(P, epi, s);
}

// tables for the fc1 epilogue: tab_act[net][a][h] = fc1_w[h][O + a];  tab_id[net][n][h] = fc1_w[h][O + A' + n] + b1[h]
__global__ void fc1_tables_kernel(const float* __restrict__ w_on, const float* __restrict__ b_on,
                                  const float* __restrict__ w_tg, const float* __restrict__ b_tg, int O, int A, int N,
                                  int D_in, int use_act, int use_id, float* __restrict__ tab_act,
                                  float* __restrict__ tab_id) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n_act = 2 * A * 64, n_id = 2 * N * 64;
    if (i < n_act) {
        int net = i / (A * 64), r = i - net * A * 64, a = r / 64, h = r - a * 64;
        const float* w = net ? w_tg : w_on;
        tab_act[i] = use_act ? w[(int64_t)h * D_in + O + a] : 0.f;
    } else if (i < n_act + n_id) {
        int k = i - n_act;
        int net = k / (N * 64), r = k - net * N * 64, n = r / 64, h = r - n * 64;
        const float* w = net ? w_tg : w_on;
        const float* b = net ? b_tg : b_on;
        tab_id[k] = b[h] + (use_id ? w[(int64_t)h * D_in + O + (use_act ? A : 0) + n] : 0.f);
    }
}

int64_t tc_fc1_scratch_bytes(const pmb_dims* d);

// streaming fc1: packed W with the agent-id columns and the bias folded into the K padding (fold_id), and the
// last-action table of both nets as bf16 [A][net][64]
__global__ void fc1_stream_pack_kernel(const float* __restrict__ w_on, const float* __restrict__ b_on,
                                       const float* __restrict__ w_tg, const float* __restrict__ b_tg, int O, int A, int N,
                                       int D_in, int use_act, int use_id, int fold_id, int n_chunks,
                                       __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ tab_act16) {
    const int64_t n_w = (int64_t)128 * n_chunks * 8;          // one thread per 16-byte chunk of the packed image
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_w) {
        const int j = (int)(i & 7);
        const int col = (int)((i >> 3) & 127);
        const int c = (int)(i >> 10);
        const int net = col >> 6, h = col & 63;
        const float* w = (net ? w_tg : w_on) + (int64_t)h * D_in;
        const float bias = (net ? b_tg : b_on)[h];
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = c * 64 + j * 8 + e;
            float v = 0.f;
            if (k < O) v = w[k];
            else if (fold_id) {
                const int q = k - O;
                if (q < N) v = use_id ? w[O + (use_act ? A : 0) + q] : 0.f;
                else if (q == N) v = bias;
            }
            f[e] = v;
        }
        char* tile = reinterpret_cast<char*>(wp) + (int64_t)c * 16384;
        *reinterpret_cast<uint4*>(tile + tc::sw128_offset((uint32_t)col, (uint32_t)j)) =
            make_uint4(tc::pack_bf16x2(f[0], f[1]), tc::pack_bf16x2(f[2], f[3]), tc::pack_bf16x2(f[4], f[5]),
                       tc::pack_bf16x2(f[6], f[7]));
        return;
    }
    i -= n_w;
    if (i < (int64_t)A * 128) {
        const int a = (int)(i >> 7), net = (int)((i >> 6) & 1), h = (int)(i & 63);
        const float* w = net ? w_tg : w_on;
        tab_act16[i] = __float2bfloat16(use_act ? w[(int64_t)h * D_in + O + a] : 0.f);
    }
}

#ifdef PMB_FC1_PROFILE
extern "C" int pmb_debug_fc1_prof(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, tc::g_fc1_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tc::g_fc1_prof, z, sizeof(z)); }
    return 0;
}
#endif
int tc_fc1_fwd_both(const pmb_dims* d, const pmb_batch* b, int t0, int nt, const AgentParams& on, const AgentParams& tg,
                    float* x_on, float* x_tg, int tile_images, uint8_t* obs_img_out, uint32_t* relu_mask, void* scratch,
                    int64_t scratch_bytes, cudaStream_t s, int weights_packed) {
    // scratch: packed W (128 x Kpad bf16) | tab_act | tab_id
    // weights_packed: the caller guarantees that `scratch` still holds the images / tables a previous call packed from
    // the SAME parameters (rollout steps between two learner updates): the pack launches are skipped
    const int D_in = d_in_of(d);
    int64_t wp_bytes = align_up(tc_packed_elems(128, d->O) * 2, 256);
    int64_t ta_bytes = align_up((int64_t)2 * d->A * 64 * 4, 256);
    int64_t ti_bytes = align_up((int64_t)2 * d->N * 64 * 4, 256);
    if (scratch_bytes < tc_fc1_scratch_bytes(d)) { set_error("tc_fc1: scratch too small"); return PMB_ERR_WORKSPACE; }
    __nv_bfloat16* wp = static_cast<__nv_bfloat16*>(scratch);
    float* tab_act = reinterpret_cast<float*>(static_cast<char*>(scratch) + wp_bytes);
    float* tab_id = reinterpret_cast<float*>(static_cast<char*>(scratch) + wp_bytes + ta_bytes);
    // the streaming kernel packs its own W (fc1_stream_pack_kernel) and needs the agent-id table only when the one-hot
    // columns do not fit the K padding
    const int n_chunks_s = (d->O + tc::BK - 1) / tc::BK;
    // a group is at most (GROUP_ROWS - 1) / N + 2 contiguous runs, each displaced by < 320 bytes (see the producer)
    const int max_runs_s = std::min(tc::fs::GROUP_ROWS, (tc::fs::GROUP_ROWS - 1) / d->N + 2);
    const int slot_bytes_s = (int)align_up((int64_t)tc::fs::GROUP_ROWS * d->O * 4 + 320 * max_runs_s + 128, 128);
    const int64_t smem_need_s = 1024 + 2 * (int64_t)n_chunks_s * tc::A_STAGE_BYTES + tc::fs::N_SLOTS * (int64_t)slot_bytes_s +
                                2 * tc::fs::N_SLOTS * tc::fs::GROUP_ROWS * 4 + 256;
    const bool use_stream = tile_images && n_chunks_s <= tc::fs::MAX_CHUNKS && smem_need_s <= 232448;
    const bool fold_id_s = n_chunks_s * tc::BK - d->O >= d->N + 1;
    int rc = PMB_OK;
    if (!use_stream) {
        const float* ptrs[2] = {on.fc1_w, tg.fc1_w};
        int rows[2] = {64, 64}, lds[2] = {D_in, D_in};
        rc = tc_pack_w(ptrs, rows, lds, 2, d->O, wp, s);
        if (rc) return rc;
    }
    if ((!use_stream || !fold_id_s) && !(weights_packed && use_stream)) {
        int n_tab = 2 * (d->A + d->N) * 64;
        fc1_tables_kernel<<<(unsigned)ceil_div(n_tab, 256), 256, 0, s>>>(on.fc1_w, on.fc1_b, tg.fc1_w, tg.fc1_b, d->O, d->A,
                                                                         d->N, D_in, d->obs_last_action, d->obs_agent_id,
                                                                         tab_act, tab_id);
        PMB_LAUNCH_CHECK("fc1_tables_kernel");
    }
    const int64_t M = (int64_t)d->B * nt * d->N;
    RowMap map{b->obs_sb, (int64_t)d->N * d->O, (int64_t)d->O, nt, d->N};
    tc::GemmParams P{b->obs + (int64_t)t0 * d->N * d->O, map, M, d->O, (d->O + tc::BK - 1) / tc::BK, wp, 128,
                     0, 0, 0, 0, nullptr, nullptr};
    if (tile_images) {
        // rows in time-major tiled order so that GEMM tiles coincide with the (t, tile) tile images
        const int64_t R = (int64_t)d->B * d->N;
        const int n_tiles = (int)ceil_div(R, 128);
        P.M = (int64_t)nt * n_tiles * 128;
        P.tm_R = R; P.tm_Rpad = (int64_t)n_tiles * 128; P.tm_N = d->N; P.tm_T = nt;
        P.a_img_out = obs_img_out;
        tc::Fc1TiEpi epi{tab_act, tab_id, b->actions, b->actions_sb, b->filled, b->filled_sb,
                         reinterpret_cast<uint8_t*>(x_on), reinterpret_cast<uint8_t*>(x_tg), t0, nt, d->N, d->A,
                         d->obs_last_action, n_tiles, R, relu_mask};
        const int n_chunks = (d->O + tc::BK - 1) / tc::BK;
        const int slot_bytes = slot_bytes_s;
        const int64_t smem_need = 1024 + 2 * (int64_t)n_chunks * tc::A_STAGE_BYTES + tc::fs::N_SLOTS * (int64_t)slot_bytes +
                                  2 * tc::fs::N_SLOTS * tc::fs::GROUP_ROWS * 4 + 256;
        if (n_chunks <= tc::fs::MAX_CHUNKS && smem_need <= 232448) {
            const int fold_id = n_chunks * tc::BK - d->O >= d->N + 1;
            __nv_bfloat16* tab_act16 = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(scratch) + wp_bytes + ta_bytes + ti_bytes);
            const int64_t n_pack = (int64_t)128 * n_chunks * 8 + (int64_t)d->A * 128;
            if (!weights_packed) {
                fc1_stream_pack_kernel<<<(unsigned)ceil_div(n_pack, 256), 256, 0, s>>>(
                    on.fc1_w, on.fc1_b, tg.fc1_w, tg.fc1_b, d->O, d->A, d->N, D_in, d->obs_last_action, d->obs_agent_id, fold_id,
                    n_chunks, wp, tab_act16);
                PMB_LAUNCH_CHECK("fc1_stream_pack_kernel");
            }
            tc::Fc1StreamParams Q;
            Q.ep_index = b->ep_index; Q.obs = b->obs; Q.obs_sb = b->obs_sb; Q.Wp = wp; Q.obs_img = obs_img_out; Q.R = R;
            Q.T = nt; Q.N = d->N; Q.O = d->O; Q.n_tiles = n_tiles; Q.n_chunks = n_chunks; Q.slot_bytes = slot_bytes;
            Q.fold_id = fold_id; Q.tab_act16 = reinterpret_cast<const uint4*>(tab_act16); Q.tab_id = tab_id;
            Q.actions = b->actions; Q.actions_sb = b->actions_sb; Q.filled = b->filled; Q.filled_sb = b->filled_sb;
            Q.x_on = reinterpret_cast<uint8_t*>(x_on); Q.x_tg = reinterpret_cast<uint8_t*>(x_tg);
            Q.relu_mask = relu_mask; Q.use_act = d->obs_last_action;
            Q.n_nets = x_tg ? 2 : 1; Q.t0 = t0;
            // L2 prefetch of the CTA's next tile(s): measured (round 2, same box, alternating) as a LOSS - rollout step 0.196 ms
            // with one tile ahead, 0.217 ms with two, 0.190 ms without; learner fc1 8.1-8.4 / 8.8 / 8.0 ms - the copies of the
            // ring already keep the memory system busy and the prefetch instructions cost the producers issue time.
            // PMB_FC1_PREFETCH = tiles ahead keeps the experiment available.
            Q.l2_prefetch = 0;
            { const char* e = getenv("PMB_FC1_PREFETCH"); if (e) Q.l2_prefetch = atoi(e); }
            const int64_t n_items = (int64_t)nt * n_tiles;
            const int grid = (int)(n_items < sm_count() ? n_items : sm_count());
            // one net (rollout step): W takes half the shared memory, the staging ring gets two more slots when they fit -
            // the stream is bound by the bytes in flight per SM (slot turn-around ~3500 cycles: 3 x 18 KB give 4.6 TB/s)
            const int64_t smem_1net = 1024 + (int64_t)n_chunks * (tc::A_STAGE_BYTES / 2) + (int64_t)n_chunks * tc::A_STAGE_BYTES +
                                      tc::fs::N_SLOTS_1NET * (int64_t)slot_bytes + 2 * tc::fs::N_SLOTS_1NET * tc::fs::GROUP_ROWS * 4 + 256;
            const bool wide_ring = Q.n_nets == 1 && smem_1net <= 232448;
#define PMB_FC1_STREAM_LAUNCH(NC)                                                                                            \
    case NC:                                                                                                                 \
        if (wide_ring) {                                                                                                     \
            PMB_SMEM_ATTR((tc::fc1_stream_kernel<NC, tc::fs::N_SLOTS_1NET>), (int)smem_1net);                               \
            PMB_CUDA(launch_pdl(tc::fc1_stream_kernel<NC, tc::fs::N_SLOTS_1NET>, dim3(grid),                                 \
                                dim3(tc::fs::threads_for(tc::fs::N_SLOTS_1NET)), (size_t)smem_1net, s, rollout_pdl_enabled(), Q)); \
        } else {                                                                                                             \
            PMB_SMEM_ATTR((tc::fc1_stream_kernel<NC, tc::fs::N_SLOTS>), (int)smem_need);                                     \
            tc::fc1_stream_kernel<NC, tc::fs::N_SLOTS><<<grid, tc::fs::threads_for(tc::fs::N_SLOTS), (size_t)smem_need, s>>>(Q); \
        }                                                                                                                    \
        break;
            switch (n_chunks) {
                PMB_FC1_STREAM_LAUNCH(1)
                PMB_FC1_STREAM_LAUNCH(2)
                PMB_FC1_STREAM_LAUNCH(3)
                PMB_FC1_STREAM_LAUNCH(4)
                PMB_FC1_STREAM_LAUNCH(5)
            }
#undef PMB_FC1_STREAM_LAUNCH
            PMB_LAUNCH_CHECK("fc1_stream_kernel");
            return PMB_OK;
        }
        if (b->ep_index) { set_error("tc_fc1: ep_index needs the streaming kernel (obs dim <= 320)"); return PMB_ERR_INVALID; }
        return tc::launch_tc_gemm(P, epi, s);
    }
    if (b->ep_index) { set_error("tc_fc1: ep_index is not supported without tile images"); return PMB_ERR_INVALID; }
    tc::Fc1Epi epi{tab_act, tab_id, b->actions, b->actions_sb, b->filled, b->filled_sb, x_on, x_tg,
                   t0, nt, d->N, d->A, d->obs_last_action, (int64_t)d->B * d->N};
    return tc::launch_tc_gemm(P, epi, s);
}
int64_t tc_fc1_scratch_bytes(const pmb_dims* d) {
    return align_up(tc_packed_elems(128, d->O) * 2, 256) + align_up((int64_t)2 * d->A * 64 * 4, 256) +
           align_up((int64_t)2 * d->N * 64 * 4, 256) + align_up((int64_t)d->A * 128 * 2, 256);
}

// permuted bias for the mixer: packed order [w1 | b1 | w_final | v0] from flat order [w1 | w_final | b1 | v0]
__global__ void mix_bias_perm_kernel(const float* __restrict__ b_cat, int N, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int C = (N + 3) * 32;
    if (i >= C) return;
    int grp = i >> 5, e = i & 31;
    int src_grp = grp < N ? grp : (grp == N ? N + 1 : (grp == N + 1 ? N : N + 2));
    out[i] = b_cat[src_grp * 32 + e];
}

int64_t tc_mixer_scratch_bytes(const pmb_dims* d) {
    const int C = (d->N + 3) * 32;
    return align_up(tc_packed_elems((int)align_up(C, 32), d->S + 1) * 2, 256) + align_up((int64_t)C * 4, 256);
}

// ---- image-fed mixer path -------------------------------------------------------------------------
// state [B, T, S] fp32 -> bf16 tile images [ceil(B*T/128)][ceil((S+1)/64)][16 KB] over ALL (b, t) rows, m' = b*T + t.
// Column S of every real row holds 1.0 (the bias-gradient column of the hypernet weight-gradient GEMM; the packed
// forward weights are zero there), everything else beyond S and all padding rows are zero.
__global__ void __launch_bounds__(256)
state_to_images_kernel(const float* __restrict__ state, int64_t state_sb, const int64_t* __restrict__ ep_index, int64_t BT,
                       int T, int S, int n_chunks, uint8_t* __restrict__ img) {
    // one thread per 16-byte chunk
    const int64_t total = ((BT + 127) / 128) * 128 * (int64_t)n_chunks * 8;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int j = (int)(i & 7);
    int64_t q = i >> 3;
    const int r = (int)(q & 127);
    q >>= 7;
    const int c = (int)(q % n_chunks);
    const int64_t tile = q / n_chunks;
    const int64_t m = tile * 128 + r;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (m < BT) {
        const int64_t b = m / T;
        const int t = (int)(m - b * T);
        const float* src = state + ep_row(ep_index, b) * state_sb + (int64_t)t * S + c * 64 + j * 8;
        const int nv = S - (c * 64 + j * 8);
        if (nv >= 8 && (reinterpret_cast<uintptr_t>(src) & 7) == 0) {        // the common case: four 8-byte loads
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(src) + e);
                f[2 * e] = v.x; f[2 * e + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (e < nv) f[e] = __ldg(src + e);
                else if (e == nv) f[e] = 1.0f;
            }
        }
    }
    *reinterpret_cast<uint4*>(img + (tile * n_chunks + c) * 16384 + tc::sw128_offset((uint32_t)r, (uint32_t)j)) =
        make_uint4(tc::pack_bf16x2(f[0], f[1]), tc::pack_bf16x2(f[2], f[3]), tc::pack_bf16x2(f[4], f[5]),
                   tc::pack_bf16x2(f[6], f[7]));
}

int tc_state_to_images(const pmb_dims* d, const pmb_batch* b, uint8_t* img, cudaStream_t s) {
    const int64_t BT = (int64_t)d->B * d->T;
    const int n_chunks = tc_state_chunks(d);
    const int64_t total = ((BT + 127) / 128) * 128 * (int64_t)n_chunks * 8;
    state_to_images_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(b->state, b->state_sb, b->ep_index, BT, d->T, d->S, n_chunks,
                                                                                  img);
    PMB_LAUNCH_CHECK("state_to_images_kernel");
    return PMB_OK;
}

// QMIX forward from state images; raw_img != null keeps the hypernet outputs (online mixer)
int tc_mixer_fwd_img(const pmb_dims* d, const MixerParams& mp, const uint8_t* state_img, const float* agent_qs, int t_off,
                     uint8_t* raw_img, float* q_tot, void* scratch, int64_t scratch_bytes, cudaStream_t s) {
    const int N = d->N, S = d->S, C = (N + 3) * 32;
    const int n_chunks = tc_state_chunks(d);
    if (scratch_bytes < tc_mixer_scratch_bytes(d)) { set_error("tc_mixer: scratch too small"); return PMB_ERR_WORKSPACE; }
    __nv_bfloat16* wp = static_cast<__nv_bfloat16*>(scratch);
    float* bias = reinterpret_cast<float*>(static_cast<char*>(scratch) + align_up(tc_packed_elems(C, S + 1) * 2, 256));
    const float* ptrs[4] = {mp.w_cat, mp.w_cat + (int64_t)(N + 1) * 32 * S, mp.w_cat + (int64_t)N * 32 * S,
                            mp.w_cat + (int64_t)(N + 2) * 32 * S};
    int rows[4] = {N * 32, 32, 32, 32}, lds[4] = {S, S, S, S};
    int rc = tc_pack_w(ptrs, rows, lds, 4, S, wp, s, n_chunks);
    if (rc) return rc;
    mix_bias_perm_kernel<<<(unsigned)ceil_div(C, 256), 256, 0, s>>>(mp.b_cat, N, bias);
    PMB_LAUNCH_CHECK("mix_bias_perm_kernel");
    const int64_t BT = (int64_t)d->B * d->T;
    tc::GemmParams P{nullptr, dense_map(0), ((BT + 127) / 128) * 128, n_chunks * 64, n_chunks, wp, C,
                     0, 0, 0, 0, nullptr, state_img};
    tc::MixImgEpi epi{bias, agent_qs, mp.v2_w, mp.v2_b, q_tot, raw_img, N, d->T, t_off, tc_mix_cblks(d), BT};
    return tc::launch_tc_gemm(P, epi, s);
}

}  // namespace pmb
