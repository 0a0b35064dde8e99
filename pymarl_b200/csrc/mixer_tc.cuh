// Image-fed QMIX mixer of the bf16 tensor-core tier (tc_gemm.cu: forward, mixer_tc.cu: backward).
//
//   state images : [ceil(B*T/128)][tc_state_chunks][16 KB]  bf16 K-major / 128B-swizzle tile images of ALL (b, t)
//                  rows (m' = b*T + t), written once per step; column S holds 1.0 (bias-gradient column).
//   raw images   : [ceil(B*T/128)][tc_mix_cblks][16 KB]     hypernet outputs of the ONLINE mixer in the packed column
//                  order [w1 (N*32) | b1 | w_final | v0]; the backward overwrites them with d_raw in place.
// Both mixers run their GEMM over all B*T rows (1/T wasted) so that the two passes and the weight-gradient GEMM share
// one set of images; the epilogue maps row (b, t) to the mixer row b*(T-1) + t - t_off.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace pmb {

int64_t tc_packed_elems(int Ncols, int K);
int tc_pack_w(const float* const* ptrs, const int* rows, const int* lds, int nseg, int K, __nv_bfloat16* out,
              cudaStream_t s, int n_chunks_min = 0);

inline int tc_state_chunks(const pmb_dims* d) { return (d->S + 1 + 63) / 64; }
inline int tc_mix_cblks(const pmb_dims* d) { return (d->N + 3 + 1) / 2; }
inline int64_t tc_mix_row_tiles(const pmb_dims* d) { return ((int64_t)d->B * d->T + 127) / 128; }
inline int64_t tc_state_img_bytes(const pmb_dims* d) { return tc_mix_row_tiles(d) * tc_state_chunks(d) * 16384; }
inline int64_t tc_raw_img_bytes(const pmb_dims* d) { return tc_mix_row_tiles(d) * tc_mix_cblks(d) * 16384; }

int64_t tc_mixer_scratch_bytes(const pmb_dims* d);
int tc_state_to_images(const pmb_dims* d, const pmb_batch* b, uint8_t* img, cudaStream_t s);
int tc_mixer_fwd_img(const pmb_dims* d, const MixerParams& mp, const uint8_t* state_img, const float* agent_qs, int t_off,
                     uint8_t* raw_img, float* q_tot, void* scratch, int64_t scratch_bytes, cudaStream_t s);

// backward: d_raw in place, d_agent_qs, V.2 gradients, then hypernet weight / bias gradients = d_raw^T . [state | 1]
int64_t tc_mixer_bwd_img_scratch_bytes(const pmb_dims* d);
// the two halves of tc_mixer_bwd_img, so that the weight-gradient GEMM can run on a side stream next to the agent's BPTT
int tc_mixer_bwd_img_dq(const pmb_dims* d, const MixerParams& mp, uint8_t* raw_img, const float* agent_qs, const float* g,
                        float* d_agent_qs, float* gv2_w, float* gv2_b, void* scratch, int64_t scratch_bytes, cudaStream_t s);
int tc_mixer_dw_ctas(const pmb_dims* d);                 // CTAs of the weight-gradient GEMM with one row slice
int64_t tc_mixer_dw_scratch_bytes(const pmb_dims* d);
int tc_mixer_dw(const pmb_dims* d, const uint8_t* state_img, const uint8_t* raw_img, float* gw_cat, float* gb_cat,
                void* scratch, int64_t scratch_bytes, int ctas_avail, cudaStream_t s);
int tc_mixer_bwd_img(const pmb_dims* d, const MixerParams& mp, const uint8_t* state_img, uint8_t* raw_img,
                     const float* agent_qs, const float* g, float* d_agent_qs, float* gw_cat, float* gb_cat, float* gv2_w,
                     float* gv2_b, void* scratch, int64_t scratch_bytes, cudaStream_t s);

}  // namespace pmb
