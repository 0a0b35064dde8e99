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
