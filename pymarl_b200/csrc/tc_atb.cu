// bf16 tensor-core tier, weight-gradient GEMM:  out[c, k] = sum_m D[m, c] * A[m, k]   (reduction over rows)
//
// Both operands are "MN-major" for the tensor core (the reduction index m is the slow memory index), so
// the shared-memory tiles are built in the MN-major 128-byte-swizzle canonical layout and the instruction
// descriptor sets a_major = b_major = MN.  Sources are fp32 in global memory (activations / batch fields,
// addressed through RowMap); producer warps convert to bf16 on the way into shared memory.
//   tile:   128 (c) x <=256 (k) fp32 accumulator in TMEM, 64 rows of m per pipeline stage (4 UMMA K=16)
//   split:  grid.z slices of the row range, each writes its partial tile; a fixed-order reduction
//           (reduce_slices) sums them -> deterministic, no float atomics.
//   bias:   column `K` of the logical A operand is all ones, so out[c, K] = sum_m D[m, c] (the bias
//           gradient) comes out of the same MMA.
// Warp roles (544 threads): 0-3 epilogue, 4 TMEM alloc + MMA issuer, 5-8 D producers, 9-16 A producers.
#include "common.cuh"
#include "tc_common.cuh"

namespace pmb {
namespace tc {

namespace atb {
constexpr int BC = 128;              // UMMA M (output rows = D columns)
constexpr int BKN = 256;             // UMMA N max (output cols = A columns)
constexpr int BR = 64;               // rows of m per stage
constexpr int STAGES = 4;
constexpr int D_STAGE_BYTES = BC * BR * 2;        // 16 KB : 2 blocks of (64 rows x 128 B)
constexpr int A_STAGE_BYTES = BKN * BR * 2;       // 32 KB : 4 blocks
constexpr int BLOCK_BYTES = BR * 128;             // one 64-element MN block: 64 rows x 128 B
constexpr int MMA_WARP = 4, FIRST_D_WARP = 5, N_D_WARPS = 4, FIRST_A_WARP = 9, N_A_WARPS = 8;
constexpr int THREADS = 32 * (FIRST_A_WARP + N_A_WARPS);       // 544
constexpr int SMEM_BYTES = 1024 + STAGES * (D_STAGE_BYTES + A_STAGE_BYTES) + 256;
}  // namespace atb

struct AtbParams {
    const float* D; RowMap dmap; int C;
    const float* A; RowMap amap; int K;          // logical A has K + 1 columns (last = ones)
    int64_t M;
    int64_t rows_per_slice;                       // multiple of 64
    float* partial;                               // [slices][C][K + 1]
    int tile_w;                                   // columns per n-tile (multiple of 16, <= 256), balanced
    // tile-image mode (D = bf16 [T][n_tiles][16 KB] images, C = 64): rows are m = t * R_pad + p
    const uint8_t* d_ti;
    int ti_T, ti_N, ti_n_tiles;
    int64_t ti_R;
};

// load 4 consecutive floats p[0..3] with per-element validity, using the widest aligned access
__device__ __forceinline__ void load4_masked(const float* p, int n_valid, float (&v)[4]) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (n_valid >= 4) {
        uintptr_t a = reinterpret_cast<uintptr_t>(p);
        if ((a & 15) == 0) {
            float4 t = __ldg(reinterpret_cast<const float4*>(p));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else if ((a & 7) == 0) {
            float2 t0 = __ldg(reinterpret_cast<const float2*>(p));
            float2 t1 = __ldg(reinterpret_cast<const float2*>(p + 2));
            v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
        } else {
            v[0] = __ldg(p); v[1] = __ldg(p + 1); v[2] = __ldg(p + 2); v[3] = __ldg(p + 3);
        }
    } else {
        if (n_valid > 0) v[0] = __ldg(p);
        if (n_valid > 1) v[1] = __ldg(p + 1);
        if (n_valid > 2) v[2] = __ldg(p + 2);
    }
}

// MN-major tile: element (mn, kr) lives in block mn/64, row kr, 16-byte chunk (mn%64)/8, swizzled by kr%8
__device__ __forceinline__ uint32_t mn_tile_offset(int mn, int kr) {
    return (uint32_t)(mn >> 6) * atb::BLOCK_BYTES + sw128_offset((uint32_t)kr, (uint32_t)((mn & 63) >> 3)) +
           (uint32_t)(mn & 7) * 2u;
}

__global__ void __launch_bounds__(atb::THREADS, 1) tc_atb_kernel(AtbParams P) {
    using namespace atb;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* d_stage = smem;
    uint8_t* a_stage = smem + STAGES * D_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (D_STAGE_BYTES + A_STAGE_BYTES));
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K1 = P.K + 1;                                     // logical columns incl. the ones column
    const int c0 = blockIdx.y * BC;
    const int k0 = blockIdx.x * P.tile_w;
    int ncols = K1 - k0 < P.tile_w ? K1 - k0 : P.tile_w;        // real columns of this tile
    const int ncols_pad = ncols <= 16 ? 16 : ((ncols + 15) & ~15);
    const int64_t mbeg = (int64_t)blockIdx.z * P.rows_per_slice;
    const int64_t mend = mbeg + P.rows_per_slice < P.M ? mbeg + P.rows_per_slice : P.M;
    const int n_chunks = mend > mbeg ? (int)((mend - mbeg + BR - 1) / BR) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], N_D_WARPS + N_A_WARPS); mbar_init(&empty[s], 1); }
        mbar_init(&tfull[0], 1);
        fence_barrier_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, 256);
    if (P.d_ti) {                                               // MN block 1 of every D stage stays zero
        for (int i = threadIdx.x; i < STAGES * (BLOCK_BYTES / 16); i += THREADS) {
            const int s = i / (BLOCK_BYTES / 16), q = i - s * (BLOCK_BYTES / 16);
            reinterpret_cast<uint4*>(d_stage + s * D_STAGE_BYTES + BLOCK_BYTES)[q] = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= FIRST_A_WARP) {
        // ===== A producers: rows of A (fp32) -> MN-major bf16 tile, + the ones column =====
        // All loads of a chunk are issued before the first conversion (memory-level parallelism:
        // the kernel is HBM/latency bound, not tensor bound).
        constexpr int RPW = BR / N_A_WARPS;                       // 8 rows per warp per chunk
        const int pw = warp - FIRST_A_WARP;
        const int n_it = (ncols_pad + 127) / 128;                 // 1 or 2 column sweeps of 128
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int s = ch % STAGES;
            // lane r (< RPW) resolves the row pointer of tile row (pw + N_A_WARPS r)
            const float* myptr = nullptr;
            if (lane < RPW) {
                const int64_t m = mbeg + (int64_t)ch * BR + pw + N_A_WARPS * lane;
                if (m < mend) {
                    if (P.d_ti) {
                        const int64_t r_pad = (int64_t)P.ti_n_tiles * 128;
                        const int64_t t = m / r_pad, pp = m - t * r_pad;
                        if (pp < P.ti_R) {
                            const int64_t bb = pp / P.ti_N, nn = pp - bb * P.ti_N;
                            myptr = P.A + P.amap.offset((bb * P.ti_T + t) * P.ti_N + nn);
                        }
                    } else {
                        myptr = P.A + P.amap.offset(m);
                    }
                }
            }
            float v[RPW][2][4];
            uint32_t row_ok = 0;
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const float* rp = reinterpret_cast<const float*>(
                    __shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(myptr), rr));
                row_ok |= (rp != nullptr ? 1u : 0u) << rr;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    v[rr][q][0] = v[rr][q][1] = v[rr][q][2] = v[rr][q][3] = 0.f;
                    const int cc = q * 128 + 4 * lane;
                    if (q < n_it && cc < ncols_pad && rp != nullptr) {
                        const int kcol = k0 + cc;
                        int nv = P.K - kcol;
                        if (nv > 0) load4_masked(rp + kcol, nv, v[rr][q]);      // consumed only after ALL loads are issued
                    }
                }
            }
            mbar_wait(&empty[s], ((ch / STAGES) & 1) ^ 1);
            uint8_t* dst = a_stage + s * A_STAGE_BYTES;
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const int kr = pw + N_A_WARPS * rr;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int cc = q * 128 + 4 * lane;
                    if (q < n_it && cc < ncols_pad) {
                        // the ones column (logical index K) of valid rows -> bias gradient
                        const int nv = P.K - (k0 + cc);
                        const float one = (row_ok >> rr) & 1u ? 1.0f : 0.0f;
                        float a0 = nv == 0 ? one : v[rr][q][0], a1 = nv == 1 ? one : v[rr][q][1];
                        float a2 = nv == 2 ? one : v[rr][q][2], a3 = nv == 3 ? one : v[rr][q][3];
                        *reinterpret_cast<uint2*>(dst + mn_tile_offset(cc, kr)) =
                            make_uint2(pack_bf16x2(a0, a1), pack_bf16x2(a2, a3));
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
    } else if (warp >= FIRST_D_WARP) {
        // ===== D producers: rows of D (fp32), 128 columns c0.. -> MN-major bf16 tile =====
        constexpr int RPW = BR / N_D_WARPS;                       // 16 rows per warp per chunk
        const int pw = warp - FIRST_D_WARP;
        const int cc = 4 * lane;
        const int nv = P.C - (c0 + cc);
        if (P.d_ti) {
            // tile-image mode: the 64 rows of a chunk are one contiguous 8 KB piece of a tile image
            // (= MN block 0); block 1 of every stage was zeroed at start.  One bulk copy per chunk.
            for (int ch = 0; ch < n_chunks; ++ch) {
                const int s = ch % STAGES;
                mbar_wait(&empty[s], ((ch / STAGES) & 1) ^ 1);
                if (lane == 0) {
                    if (pw == 0) {
                        const int64_t m0 = mbeg + (int64_t)ch * BR;
                        const int64_t r_pad = (int64_t)P.ti_n_tiles * 128;
                        const int64_t t = m0 / r_pad, p0 = m0 - t * r_pad;
                        const uint8_t* src = P.d_ti + (t * P.ti_n_tiles + (p0 >> 7)) * 16384 + (p0 & 127) * 128;
                        mbar_arrive_expect_tx(&full[s], BLOCK_BYTES);
                        bulk_copy_g2s(d_stage + s * D_STAGE_BYTES, src, BLOCK_BYTES, &full[s]);
                    } else {
                        mbar_arrive(&full[s]);
                    }
                }
            }
        } else
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int s = ch % STAGES;
            const float* myptr = nullptr;
            if (lane < RPW) {
                const int64_t m = mbeg + (int64_t)ch * BR + pw + N_D_WARPS * lane;
                if (m < mend) myptr = P.D + P.dmap.offset(m) + c0;
            }
            float v[RPW][4];
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const float* rp = reinterpret_cast<const float*>(
                    __shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(myptr), rr));
                v[rr][0] = v[rr][1] = v[rr][2] = v[rr][3] = 0.f;
                if (rp != nullptr && nv > 0) load4_masked(rp + cc, nv, v[rr]);
            }
            mbar_wait(&empty[s], ((ch / STAGES) & 1) ^ 1);
            uint8_t* dst = d_stage + s * D_STAGE_BYTES;
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr)
                *reinterpret_cast<uint2*>(dst + mn_tile_offset(cc, pw + N_D_WARPS * rr)) =
                    make_uint2(pack_bf16x2(v[rr][0], v[rr][1]), pack_bf16x2(v[rr][2], v[rr][3]));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0 && n_chunks > 0) {
            const uint32_t idesc = umma_idesc_bf16(BC, ncols_pad, 1, 1);
            for (int ch = 0; ch < n_chunks; ++ch) {
                const int s = ch % STAGES;
                mbar_wait(&full[s], (ch / STAGES) & 1);
                tc_fence_after();
                const uint32_t d_addr = smem_u32(d_stage + s * D_STAGE_BYTES);
                const uint32_t a_addr = smem_u32(a_stage + s * A_STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < BR / 16; ++kk) {
                    // 16 rows of m = two 8-row groups: advance by 2 * SBO
                    umma_bf16(tmem_base, umma_desc_sw128(d_addr + kk * 2048, BLOCK_BYTES, 1024),
                              umma_desc_sw128(a_addr + kk * 2048, BLOCK_BYTES, 1024), idesc, (ch | kk) != 0);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(&tfull[0]);
        }
    } else {
        // ===== epilogue: partial[slice][c][k] =====
        const int c = c0 + warp * 32 + lane;
        float* out = P.partial + ((int64_t)blockIdx.z * P.C + c) * K1 + k0;
        if (n_chunks > 0) {
            mbar_wait(&tfull[0], 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int g = 0; g < ncols_pad; g += 32) {
                uint32_t v[32];
                if (g + 32 <= ncols_pad) {
                    tmem_ld_32x32(taddr + g, v);
                } else {                                   // 16-column tail
                    uint32_t t[16];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
                          "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]),
                          "=r"(t[15])
                        : "r"(taddr + g)
                        : "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) { v[j] = t[j]; v[16 + j] = 0u; }
                }
                tmem_wait_ld();
                if (c < P.C) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (g + j < ncols) out[g + j] = __uint_as_float(v[j]);
                }
            }
        } else if (c < P.C) {
            for (int j = 0; j < ncols; ++j) out[j] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// out[c*ldo + k] = sum_s partial[s][c][k] (k < K);  bias_out[c] = sum_s partial[s][c][K]
__global__ void atb_reduce_kernel(const float* __restrict__ partial, int slices, int C, int K, float* __restrict__ out,
                                  int64_t ldo, float* __restrict__ bias_out) {
    const int K1 = K + 1;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = (int64_t)C * K1;
    if (i >= n) return;
    float s = 0.f;
    for (int sl = 0; sl < slices; ++sl) s += partial[(int64_t)sl * n + i];
    int c = (int)(i / K1), k = (int)(i - (int64_t)c * K1);
    if (k < K) out[(int64_t)c * ldo + k] = s;
    else if (bias_out) bias_out[c] = s;
}

}  // namespace tc

static int tc_atb_tile_w(int K) {       // balanced n-tiles: ceil((K+1)/n_tiles) rounded up to 16
    int n_tiles = (int)ceil_div(K + 1, tc::atb::BKN);
    return (int)align_up(ceil_div(K + 1, n_tiles), 16);
}

static int tc_atb_slices(int C, int K, int64_t M) {
    int64_t tiles = ceil_div(C, tc::atb::BC) * ceil_div(K + 1, tc_atb_tile_w(K));
    int64_t want = (int64_t)sm_count() / tiles;          // one wave: never more CTAs than SMs
    int64_t mx = ceil_div(M, 4 * tc::atb::BR);
    if (want > mx) want = mx;
    if (want < 1) want = 1;
    return (int)want;
}

int64_t tc_atb_scratch_bytes(int C, int K, int64_t M) {
    return align_up((int64_t)tc_atb_slices(C, K, M) * C * (K + 1) * 4, 256);
}

int tc_gemm_atb(const float* D, RowMap dmap, int C, const float* A, RowMap amap, int K, int64_t M, float* out,
                int64_t ldo, float* bias_out, void* scratch, int64_t scratch_bytes, cudaStream_t s) {
    if (C <= 0 || K <= 0 || M <= 0) return PMB_OK;
    const int slices = tc_atb_slices(C, K, M);
    if (tc_atb_scratch_bytes(C, K, M) > scratch_bytes) {
        set_error("tc_gemm_atb: scratch too small (%lld < %lld)", (long long)scratch_bytes,
                  (long long)tc_atb_scratch_bytes(C, K, M));
        return PMB_ERR_WORKSPACE;
    }
    tc::AtbParams P;
    P.D = D; P.dmap = dmap; P.C = C; P.A = A; P.amap = amap; P.K = K; P.M = M;
    P.rows_per_slice = align_up(ceil_div(M, slices), tc::atb::BR);
    P.partial = static_cast<float*>(scratch);
    P.tile_w = tc_atb_tile_w(K);
    P.d_ti = nullptr; P.ti_T = P.ti_N = P.ti_n_tiles = 0; P.ti_R = 0;
    PMB_SMEM_ATTR(tc::tc_atb_kernel, tc::atb::SMEM_BYTES);
    dim3 grid((unsigned)ceil_div(K + 1, P.tile_w), (unsigned)ceil_div(C, tc::atb::BC), (unsigned)slices);
    tc::tc_atb_kernel<<<grid, tc::atb::THREADS, tc::atb::SMEM_BYTES, s>>>(P);
    PMB_LAUNCH_CHECK("tc_atb_kernel");
    int64_t n = (int64_t)C * (K + 1);
    tc::atb_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(P.partial, slices, C, K, out, ldo, bias_out);
    PMB_LAUNCH_CHECK("atb_reduce_kernel");
    return PMB_OK;
}

// fc1.weight[:, :O] / fc1.bias gradient: D = dpre1 tile images (C = 64), A = obs (fp32, batch major)
int tc_gemm_atb_ti(const uint8_t* d_ti, int T, int N, int64_t R, int n_tiles, const float* A, RowMap amap, int K,
                   float* out, int64_t ldo, float* bias_out, void* scratch, int64_t scratch_bytes, cudaStream_t s) {
    const int C = 64;
    const int64_t M = (int64_t)T * n_tiles * 128;
    const int slices = tc_atb_slices(C, K, M);
    if (tc_atb_scratch_bytes(C, K, M) > scratch_bytes) { set_error("tc_gemm_atb_ti: scratch too small"); return PMB_ERR_WORKSPACE; }
    tc::AtbParams P;
    P.D = nullptr; P.dmap = dense_map(0); P.C = C; P.A = A; P.amap = amap; P.K = K; P.M = M;
    P.rows_per_slice = align_up(ceil_div(M, slices), tc::atb::BR);
    P.partial = static_cast<float*>(scratch);
    P.tile_w = tc_atb_tile_w(K);
    P.d_ti = d_ti; P.ti_T = T; P.ti_N = N; P.ti_n_tiles = n_tiles; P.ti_R = R;
    PMB_SMEM_ATTR(tc::tc_atb_kernel, tc::atb::SMEM_BYTES);
    dim3 grid((unsigned)ceil_div(K + 1, P.tile_w), 1, (unsigned)slices);
    tc::tc_atb_kernel<<<grid, tc::atb::THREADS, tc::atb::SMEM_BYTES, s>>>(P);
    PMB_LAUNCH_CHECK("tc_atb_kernel<ti>");
    int64_t n = (int64_t)C * (K + 1);
    tc::atb_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(P.partial, slices, C, K, out, ldo, bias_out);
    PMB_LAUNCH_CHECK("atb_reduce_kernel");
    return PMB_OK;
}
int64_t tc_atb_ti_scratch_bytes(int T, int n_tiles, int K) { return tc_atb_scratch_bytes(64, K, (int64_t)T * n_tiles * 128); }

}  // namespace pmb
