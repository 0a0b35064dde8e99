// Replay-buffer side of the path: ReplayBuffer.sample / EpisodeBatch.__getitem__ with an
// array of episode ids (components/episode_buffer.py:205-217,291-298) and max_t_filled (:255-256) for a buffer that
// lives in HBM.  The reference gathers every field with one advanced-indexing kernel per field; here ONE launch copies
// whole episodes (contiguous [T, ...] blocks) of all fields with 16-byte accesses: a pure HBM-bandwidth kernel
// (read + write of the sampled bytes).
#include "common.cuh"

namespace pmb {
namespace {

constexpr int GATHER_MAX_FIELDS = 16;
struct GatherArgs {
    const char* src[GATHER_MAX_FIELDS];
    char* dst[GATHER_MAX_FIELDS];
    int64_t bytes[GATHER_MAX_FIELDS];        // per episode
    int32_t vec[GATHER_MAX_FIELDS];          // 16, 8, 4 or 1: widest access the pointers and the size allow
    int32_t n_fields;
    const int64_t* ids;
    int64_t n_src;                           // episodes in the source (ids are checked against it)
};

// Block x of gridDim.x copies the contiguous span [x, x+1) * ceil(n_vec / gridDim.x) of the field: consecutive threads
// take consecutive vectors, eight independent loads in flight per thread.
template <typename V>
__device__ __forceinline__ void copy_span(const char* __restrict__ s, char* __restrict__ d, int64_t n_vec) {
    const V* sv = reinterpret_cast<const V*>(s);
    V* dv = reinterpret_cast<V*>(d);
    const int64_t span = (n_vec + gridDim.x - 1) / gridDim.x;
    const int64_t beg = (int64_t)blockIdx.x * span;
    const int64_t end = beg + span < n_vec ? beg + span : n_vec;
    int64_t i = beg + threadIdx.x;
    constexpr int U = 8;
    for (; i + (U - 1) * 256 < end; i += U * 256) {
        V v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcs(sv + i + u * 256);
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(dv + i + u * 256, v[u]);
    }
    for (; i < end; i += 256) __stcs(dv + i, __ldcs(sv + i));
}

// grid: (blocks per episode, n_ids).  Block (x, j) copies its share of episode ids[j] of every field.
__global__ void __launch_bounds__(256) gather_episodes_kernel(GatherArgs A) {
    const int64_t j = blockIdx.y;
    const int64_t id = A.ids[j];
    if (id < 0 || id >= A.n_src) return;                 // invalid id: leave the destination untouched (checked on the host too)
    for (int f = 0; f < A.n_fields; ++f) {
        const char* s = A.src[f] + id * A.bytes[f];
        char* d = A.dst[f] + j * A.bytes[f];
        switch (A.vec[f]) {
            case 16: copy_span<uint4>(s, d, A.bytes[f] / 16); break;
            case 8: copy_span<uint2>(s, d, A.bytes[f] / 8); break;
            case 4: copy_span<uint32_t>(s, d, A.bytes[f] / 4); break;
            default: copy_span<uint8_t>(s, d, A.bytes[f]); break;
        }
    }
}

// max_b sum_t filled[b, t]  (episode_buffer.py:255-256), one block per 256 episodes + atomicMax
__global__ void __launch_bounds__(256) max_t_filled_kernel(const int64_t* __restrict__ filled, int64_t B, int T, int64_t sb,
                                                           unsigned long long* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long s = 0;
    if (b < B)
        for (int t = 0; t < T; ++t) s += filled[b * sb + t];
    for (int o = 16; o > 0; o >>= 1) {
        long long v = __shfl_xor_sync(0xffffffffu, s, o);
        s = v > s ? v : s;
    }
    if ((threadIdx.x & 31) == 0 && s > 0) atomicMax(out, (unsigned long long)s);
}

// ---- EpisodeBatch.update on the device -------------------------------------------
// One launch writes every field of an `update(data, bs, ts)` call (components/episode_buffer.py:98-154): cell (i, j) of
// the dense source [nb][nt][cell] goes to episode row(i) / timestep t0 + j of the destination field; `filled` is marked;
// the OneHot preprocess (transforms.py:12-21) is fused: an int64 index cell [G][1] becomes a float32 cell [G][dim].
constexpr int UPDATE_MAX_FIELDS = 12;
struct UpdateArgs {
    const char* src[UPDATE_MAX_FIELDS];
    char* dst[UPDATE_MAX_FIELDS];
    int64_t cell[UPDATE_MAX_FIELDS];         // bytes of one source cell
    int64_t sb[UPDATE_MAX_FIELDS];           // destination batch stride (bytes)
    int64_t st[UPDATE_MAX_FIELDS];           // destination time stride (bytes; 0: episode-constant field)
    int32_t vec[UPDATE_MAX_FIELDS];
    int32_t onehot[UPDATE_MAX_FIELDS];       // > 0: fused OneHot of this width
    int32_t n_fields;
    const int64_t* b_index;                  // device ids of the nb episodes, or NULL: b0 + i * b_step
    int64_t b0, b_step, nb, n_rows, t0, nt;
    int64_t* filled; int64_t filled_sb;      // elements
};

template <typename V>
__device__ __forceinline__ void copy_cell(const char* __restrict__ s, char* __restrict__ d, int64_t n_vec) {
    const V* sv = reinterpret_cast<const V*>(s);
    V* dv = reinterpret_cast<V*>(d);
    for (int64_t i = threadIdx.x; i < n_vec; i += blockDim.x) dv[i] = sv[i];
}

// one block per (episode i, timestep j) cell
__global__ void __launch_bounds__(128) batch_update_kernel(UpdateArgs A) {
    const int64_t c = blockIdx.x;
    const int64_t i = c / A.nt, j = c - i * A.nt;
    const int64_t b = A.b_index ? A.b_index[i] : A.b0 + i * A.b_step;
    if (b < 0 || b >= A.n_rows) return;                   // checked on the host as well
    const int64_t t = A.t0 + j;
    if (A.filled && threadIdx.x == 0) A.filled[b * A.filled_sb + t] = 1;
    for (int f = 0; f < A.n_fields; ++f) {
        const char* s = A.src[f] + c * A.cell[f];
        char* d = A.dst[f] + b * A.sb[f] + t * A.st[f];
        if (A.onehot[f] > 0) {
            const int dim = A.onehot[f];
            const int64_t G = A.cell[f] / 8;              // int64 indices in the cell
            const int64_t* idx = reinterpret_cast<const int64_t*>(s);
            float* o = reinterpret_cast<float*>(d);
            for (int64_t e = threadIdx.x; e < G * dim; e += blockDim.x) {
                const int64_t g = e / dim;
                o[e] = (idx[g] == e - g * dim) ? 1.f : 0.f;
            }
        } else {
            switch (A.vec[f]) {
                case 16: copy_cell<uint4>(s, d, A.cell[f] / 16); break;
                case 8: copy_cell<uint2>(s, d, A.cell[f] / 8); break;
                case 4: copy_cell<uint32_t>(s, d, A.cell[f] / 4); break;
                default: copy_cell<uint8_t>(s, d, A.cell[f]); break;
            }
        }
        // fields are applied in call order like the reference's loop over data.items(): a later field that targets the same
        // cells (insert_episode_batch copies actions_onehot AFTER the preprocess of actions produced it) wins
        __syncthreads();
    }
}

}  // namespace
}  // namespace pmb

using namespace pmb;

extern "C" {

int pmb_gather_episodes(const pmb_gather_field* fields, int32_t n_fields, const int64_t* ep_ids, int64_t n_ids,
                        int64_t n_src_episodes, pmb_stream stream) {
    PMB_REQUIRE(fields && ep_ids && n_fields > 0 && n_fields <= GATHER_MAX_FIELDS, "gather_episodes: 1..%d fields", GATHER_MAX_FIELDS);
    PMB_REQUIRE(n_ids >= 0 && n_ids < 65536 && n_src_episodes > 0, "gather_episodes: 0 <= n_ids < 65536");
    if (n_ids == 0) return PMB_OK;
    GatherArgs A;
    A.n_fields = n_fields; A.ids = ep_ids; A.n_src = n_src_episodes;
    int64_t biggest = 0;
    for (int f = 0; f < n_fields; ++f) {
        PMB_REQUIRE(fields[f].src && fields[f].dst && fields[f].bytes_per_episode > 0, "gather_episodes: field %d is empty", f);
        A.src[f] = static_cast<const char*>(fields[f].src);
        A.dst[f] = static_cast<char*>(fields[f].dst);
        A.bytes[f] = fields[f].bytes_per_episode;
        const uintptr_t bits = reinterpret_cast<uintptr_t>(fields[f].src) | reinterpret_cast<uintptr_t>(fields[f].dst) |
                               (uintptr_t)fields[f].bytes_per_episode;
        A.vec[f] = (bits & 15) == 0 ? 16 : ((bits & 7) == 0 ? 8 : ((bits & 3) == 0 ? 4 : 1));
        if (A.bytes[f] > biggest) biggest = A.bytes[f];
    }
    // enough blocks per episode to give every thread ~32 vectors of the biggest field, at least 4 waves in total
    int64_t bx = ceil_div(biggest / 16, (int64_t)256 * 32);
    const int64_t want = ceil_div((int64_t)4 * sm_count(), n_ids);
    if (bx < want) bx = want;
    if (bx > 1024) bx = 1024;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)n_ids);
    gather_episodes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A);
    PMB_LAUNCH_CHECK("gather_episodes_kernel");
    return PMB_OK;
}

int pmb_batch_update(const pmb_update_field* fields, int32_t n_fields, const int64_t* b_index, int64_t b0, int64_t b_step,
                     int64_t nb, int64_t n_rows, int64_t t0, int64_t nt, int64_t* filled, int64_t filled_sb, pmb_stream stream) {
    PMB_REQUIRE(fields && n_fields > 0 && n_fields <= UPDATE_MAX_FIELDS, "batch_update: 1..%d fields", UPDATE_MAX_FIELDS);
    PMB_REQUIRE(nb >= 0 && nt > 0 && t0 >= 0 && n_rows > 0 && nb * nt < ((int64_t)1 << 31), "batch_update: bad index range");
    PMB_REQUIRE(b_index || (b0 >= 0 && b_step > 0 && b0 + (nb - 1) * b_step < n_rows), "batch_update: episode range outside the batch");
    if (nb == 0) return PMB_OK;
    UpdateArgs A;
    A.n_fields = n_fields; A.b_index = b_index; A.b0 = b0; A.b_step = b_step; A.nb = nb; A.n_rows = n_rows; A.t0 = t0; A.nt = nt;
    A.filled = filled; A.filled_sb = filled_sb;
    for (int f = 0; f < n_fields; ++f) {
        const pmb_update_field& u = fields[f];
        PMB_REQUIRE(u.src && u.dst && u.cell_bytes > 0, "batch_update: field %d is empty", f);
        PMB_REQUIRE(u.onehot_dim <= 0 || u.cell_bytes % 8 == 0, "batch_update: a one-hot source cell holds int64 indices");
        A.src[f] = static_cast<const char*>(u.src); A.dst[f] = static_cast<char*>(u.dst);
        A.cell[f] = u.cell_bytes; A.sb[f] = u.dst_batch_stride_bytes; A.st[f] = u.dst_time_stride_bytes;
        A.onehot[f] = u.onehot_dim;
        const uintptr_t bits = reinterpret_cast<uintptr_t>(u.src) | reinterpret_cast<uintptr_t>(u.dst) | (uintptr_t)u.cell_bytes |
                               (uintptr_t)u.dst_batch_stride_bytes | (uintptr_t)u.dst_time_stride_bytes;
        A.vec[f] = (bits & 15) == 0 ? 16 : ((bits & 7) == 0 ? 8 : ((bits & 3) == 0 ? 4 : 1));
    }
    batch_update_kernel<<<(unsigned)(nb * nt), 128, 0, (cudaStream_t)stream>>>(A);
    PMB_LAUNCH_CHECK("batch_update_kernel");
    return PMB_OK;
}

int pmb_max_t_filled(const int64_t* filled, int64_t B, int32_t T, int64_t filled_sb, int64_t* out, pmb_stream stream) {
    PMB_REQUIRE(filled && out && B > 0 && T > 0, "max_t_filled: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    PMB_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), s));
    max_t_filled_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, s>>>(filled, B, T, filled_sb,
                                                                  reinterpret_cast<unsigned long long*>(out));
    PMB_LAUNCH_CHECK("max_t_filled_kernel");
    return PMB_OK;
}

}  // extern "C"
