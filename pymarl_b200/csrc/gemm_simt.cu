// fp32 CUDA-core GEMMs of the fp32 parity tier (FFMA, fp32 accumulate).
//
//   gemm_tn  : C[m, n] = sum_k A[m, k] * W[n, k]     (activations x nn.Linear weight^T)
//              128 x 64 tile, BK = 16, 256 threads, 8 x 4 outputs per thread, register
//              prefetch of the next k-tile, operands staged k-major in shared memory.
//   gemm_atb : out[c, k] = sum_m D[m, c] * A[m, k]   (weight gradients: reduction over rows)
//              64 x 64 output tile, rows streamed 16 at a time, deterministic split over m
//              (fixed slices, partials reduced in slice order - no float atomics).
//
// Rows are addressed through RowMap so the batch-major EpisodeBatch fields (obs, state) are
// consumed in place, strided batch dimension included.
#include "common.cuh"

namespace pmb {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

struct PlainEpilogue {
    const float* bias; float* C; int64_t ldc; int relu; int Ncols;
    __device__ __forceinline__ void operator()(int64_t m, int n0, const float (&v)[4]) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + j;
            if (n < Ncols) {
                float r = v[j] + (bias ? bias[n] : 0.f);
                if (relu) r = fmaxf(r, 0.f);
                C[m * ldc + n] = r;
            }
        }
    }
};

struct Fc1EpilogueDev {
    Fc1Epilogue e;
    __device__ __forceinline__ void operator()(int64_t m, int n0, const float (&v)[4]) const {
        const int tn = e.nt * e.N;
        int64_t b = m / tn;
        int r = (int)(m - b * tn);
        int tl = r / e.N;
        int n = r - tl * e.N;
        int t = e.t0 + tl;
        int a_prev = -1;
        if (e.use_act && t > 0 && e.filled[b * e.filled_sb + (t - 1)] != 0)
            a_prev = (int)e.actions[b * e.actions_sb + (int64_t)(t - 1) * e.N + n];
        float* out = e.x_out + ((int64_t)tl * e.R + b * e.N + n) * e.H;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int h = n0 + j;
            if (h < e.H) {
                const float* wrow = e.fc1_w + (int64_t)h * e.D_in;
                float x = v[j];
                if (a_prev >= 0) x += wrow[e.O + a_prev];
                if (e.use_id) x += wrow[e.O + (e.use_act ? e.A : 0) + n];
                x += e.fc1_b[h];
                out[h] = fmaxf(x, 0.f);
            }
        }
    }
};

template <class Epi>
__global__ void __launch_bounds__(NT) gemm_tn_kernel(const float* __restrict__ A, RowMap map, int64_t M, int K,
                                                     const float* __restrict__ W, int ldw, int Ncols, Epi epi) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // global -> register staging assignments
    const int a_row = tid >> 1, a_kq = (tid & 1) * 8;
    const int w_row = tid >> 2, w_kq = (tid & 3) * 4;
    const bool a_ok = (m0 + a_row) < M;
    const float* a_ptr = a_ok ? (A + map.offset(m0 + a_row)) : A;
    const bool w_ok = (n0 + w_row) < Ncols;
    const float* w_ptr = W + (int64_t)(w_ok ? (n0 + w_row) : 0) * ldw;

    float a_reg[8], w_reg[4];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int k = k0 + a_kq + i;
            a_reg[i] = (a_ok && k < K) ? __ldg(a_ptr + k) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int k = k0 + w_kq + i;
            w_reg[i] = (w_ok && k < K) ? __ldg(w_ptr + k) : 0.f;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[a_kq + i][a_row] = a_reg[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) Ws[w_kq + i][w_row] = w_reg[i];
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < K) load_tiles(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int64_t m = m0 + ty * 8 + i;
        if (m < M) epi(m, n0 + tx * 4, acc[i]);
    }
}

// ------------------------------------------------------------------------------------------
constexpr int TC = 64, TK = 64, TM = 16;

__global__ void __launch_bounds__(NT) gemm_atb_kernel(const float* __restrict__ D, RowMap dmap, int C,
                                                      const float* __restrict__ A, RowMap amap, int K, int64_t M,
                                                      int64_t rows_per_slice, float* __restrict__ partial,
                                                      float* __restrict__ bias_partial) {
    __shared__ __align__(16) float Ds[TM][TC + 4];
    __shared__ __align__(16) float As[TM][TK + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int k0 = blockIdx.x * TK, c0 = blockIdx.y * TC;
    const int64_t mbeg = (int64_t)blockIdx.z * rows_per_slice;
    const int64_t mend = mbeg + rows_per_slice < M ? mbeg + rows_per_slice : M;

    const int l_row = tid >> 4, l_q = (tid & 15) * 4;
    float acc[4][4], bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float d_reg[4], a_reg[4];
    auto load_tiles = [&](int64_t mb) {
        int64_t m = mb + l_row;
        bool ok = m < mend;
        const float* dp = D + (ok ? dmap.offset(m) : 0);
        const float* ap = A + (ok ? amap.offset(m) : 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int c = c0 + l_q + i, k = k0 + l_q + i;
            d_reg[i] = (ok && c < C) ? __ldg(dp + c) : 0.f;
            a_reg[i] = (ok && k < K) ? __ldg(ap + k) : 0.f;
        }
    };
    load_tiles(mbeg);
    for (int64_t mb = mbeg; mb < mend; mb += TM) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { Ds[l_row][l_q + i] = d_reg[i]; As[l_row][l_q + i] = a_reg[i]; }
        __syncthreads();
        if (mb + TM < mend) load_tiles(mb + TM);
#pragma unroll
        for (int mm = 0; mm < TM; ++mm) {
            float4 dv = *reinterpret_cast<const float4*>(&Ds[mm][ty * 4]);
            float4 av = *reinterpret_cast<const float4*>(&As[mm][tx * 4]);
            const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
            const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bsum[i] += d4[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d4[i], a4[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    float* P = partial + (int64_t)blockIdx.z * C * K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = c0 + ty * 4 + i;
        if (c >= C) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k < K) P[(int64_t)c * K + k] = acc[i][j];
        }
        if (bias_partial && blockIdx.x == 0 && tx == 0) bias_partial[(int64_t)blockIdx.z * C + c] = bsum[i];
    }
}

// out[c*ldo + k] = sum_s partial[s][c][k]  (slice order -> deterministic)
__global__ void reduce_slices_kernel(const float* __restrict__ partial, int slices, int C, int K, float* __restrict__ out,
                                     int64_t ldo, const float* __restrict__ bias_partial, float* __restrict__ bias_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = (int64_t)C * K;
    if (i < n) {
        float s = 0.f;
        for (int sl = 0; sl < slices; ++sl) s += partial[(int64_t)sl * n + i];
        int c = (int)(i / K), k = (int)(i - (int64_t)c * K);
        out[(int64_t)c * ldo + k] = s;
    }
    if (bias_out && i < C) {
        float s = 0.f;
        for (int sl = 0; sl < slices; ++sl) s += bias_partial[(int64_t)sl * C + i];
        bias_out[i] = s;
    }
}

int atb_slices(int C, int K, int64_t M) {
    int64_t tiles = ceil_div(C, TC) * ceil_div(K, TK);
    int64_t want = ceil_div((int64_t)4 * sm_count(), tiles);
    int64_t max_slices = ceil_div(M, 8 * TM);             // at least 128 rows per slice
    if (want > max_slices) want = max_slices;
    if (want < 1) want = 1;
    if (want > 1024) want = 1024;
    return (int)want;
}

}  // namespace

int launch_gemm_tn(const float* A, RowMap map, int64_t M, int32_t K, const float* W, int32_t ldw, int32_t Ncols,
                   const float* bias, float* C, int64_t ldc, int relu, cudaStream_t s) {
    if (M <= 0 || Ncols <= 0) return PMB_OK;
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(Ncols, BN));
    PlainEpilogue ep{bias, C, ldc, relu, Ncols};
    gemm_tn_kernel<PlainEpilogue><<<grid, NT, 0, s>>>(A, map, M, K, W, ldw, Ncols, ep);
    PMB_LAUNCH_CHECK("gemm_tn_kernel");
    return PMB_OK;
}

int launch_fc1_gemm(const float* obs, RowMap map, int64_t M, int32_t K, const float* W, int32_t ldw, int32_t Ncols,
                    const Fc1Epilogue& ep, cudaStream_t s) {
    if (M <= 0) return PMB_OK;
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(Ncols, BN));
    Fc1EpilogueDev e{ep};
    gemm_tn_kernel<Fc1EpilogueDev><<<grid, NT, 0, s>>>(obs, map, M, K, W, ldw, Ncols, e);
    PMB_LAUNCH_CHECK("gemm_tn_kernel<fc1>");
    return PMB_OK;
}

int64_t gemm_atb_scratch_bytes(int32_t C, int32_t K, int64_t M) {
    int sl = atb_slices(C, K, M);
    return align_up(((int64_t)sl * C * K + (int64_t)sl * C) * (int64_t)sizeof(float), 256);
}

int launch_gemm_atb(const float* D, RowMap dmap, int32_t C, const float* A, RowMap amap, int32_t K, int64_t M,
                    float* out, int64_t ldo, float* bias_out, void* scratch, int64_t scratch_bytes, cudaStream_t s) {
    if (C <= 0 || K <= 0) return PMB_OK;
    int sl = atb_slices(C, K, M);
    if (gemm_atb_scratch_bytes(C, K, M) > scratch_bytes) {
        set_error("gemm_atb: scratch too small (%lld < %lld)", (long long)scratch_bytes,
                  (long long)gemm_atb_scratch_bytes(C, K, M));
        return PMB_ERR_WORKSPACE;
    }
    float* partial = static_cast<float*>(scratch);
    float* bias_partial = partial + (int64_t)sl * C * K;
    int64_t rows_per_slice = align_up(ceil_div(M, sl), TM);
    dim3 grid((unsigned)ceil_div(K, TK), (unsigned)ceil_div(C, TC), (unsigned)sl);
    gemm_atb_kernel<<<grid, NT, 0, s>>>(D, dmap, C, A, amap, K, M, rows_per_slice, partial,
                                        bias_out ? bias_partial : nullptr);
    PMB_LAUNCH_CHECK("gemm_atb_kernel");
    int64_t n = (int64_t)C * K;
    reduce_slices_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(partial, sl, C, K, out, ldo, bias_partial, bias_out);
    PMB_LAUNCH_CHECK("reduce_slices_kernel");
    return PMB_OK;
}

}  // namespace pmb
