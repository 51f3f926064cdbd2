/* libp2p_b200.so -- C ABI of the B200-native Patch2Pix correlate-and-refine path.
 *
 * The reference (GrumpyZhou/patch2pix) has no FFI / operator registry: its boundary for this path
 * is the Python method surface of `networks.patch2pix.Patch2Pix`.  Each entry point
 * below therefore cites the reference method / function (file:line in the reference repo) whose
 * computation it replaces; `patch2pix_b200/model.py` binds them through ctypes behind the same
 * method names, and INTEGRATION.md shows the stub a maintainer would add to the reference.
 *
 * Conventions: every function returns 0 on success, a negative code on failure
 * (-1 invalid argument, -2 CUDA/driver error, -3 out of memory) and never throws;
 * `p2p_last_error()` returns a thread-local description of the last failure.
 * All data pointers are DEVICE pointers to caller-owned, contiguous memory on the handle's
 * device unless marked HOST.  Work is enqueued on `stream` (a cudaStream_t passed as void*);
 * no entry point synchronises the device.  The library owns only packed weights and scratch,
 * both freed by `p2p_destroy`.  There is no CPU fallback: shape violations are errors.
 */
#ifndef P2P_B200_H_
#define P2P_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2P_API __attribute__((visibility("default")))

typedef struct p2p_handle_s* p2p_handle_t;

/* HOST pointers to fp32 arrays laid out exactly as in the reference `state_dict`
 * (networks/modules.py:76-99): conv.0.weight [512,518,3,3]; conv.1.{weight,bias,running_mean,
 * running_var} [512]; conv.2.weight [512,512,3,3]; conv.3.* [512]; fc.0.weight [512,512],
 * fc.0.bias; fc.1.* [512]; fc.3.weight [256,512], fc.3.bias; fc.4.* [256]; fc.6.weight [5,256],
 * fc.6.bias [5].  BatchNorm is folded (eval mode, eps) while packing. */
typedef struct p2p_bn_s {
  const float* weight;
  const float* bias;
  const float* running_mean;
  const float* running_var;
} p2p_bn_t;

typedef struct p2p_regressor_weights_s {
  const float* conv0_weight;
  p2p_bn_t conv1_bn;
  const float* conv2_weight;
  p2p_bn_t conv3_bn;
  const float* fc0_weight;
  const float* fc0_bias;
  p2p_bn_t fc1_bn;
  const float* fc3_weight;
  const float* fc3_bias;
  p2p_bn_t fc4_bn;
  const float* fc6_weight;
  const float* fc6_bias;
  float bn_eps;
} p2p_regressor_weights_t;

P2P_API const char* p2p_last_error(void);
P2P_API int p2p_version(void);

/* One handle per device; not thread-safe; replaces the module state built by
 * Patch2Pix.__init__ (networks/patch2pix.py:13-61) for the hot path. */
P2P_API int p2p_create(int device, p2p_handle_t* out);
P2P_API int p2p_destroy(p2p_handle_t h);

/* NeighConsensus parameters, HOST fp32, reference layout (networks/ncn/conv4d.py:118-120):
 * w1 = ncn.conv.0.weight [3,16,1,3,3,3], b1 [16], w2 = ncn.conv.2.weight [3,1,16,3,3,3], b2 [1]. */
P2P_API int p2p_set_ncn_weights(p2p_handle_t h, const float* w1, const float* b1, const float* w2, const float* b2);

/* which: 0 = regress_mid, 1 = regress_fine (networks/patch2pix.py:53-58). */
P2P_API int p2p_set_regressor_weights(p2p_handle_t h, int which, const p2p_regressor_weights_t* w);

/* Options: "mid_passes"/"fine_passes" (1 = fp16 operands, 3 = fp16 hi/lo split, fp32-grade),
 * "corr_passes" (0 = CUDA-core fp32 correlation, 1/3 = tensor-core), "seg_len" (k-steps per
 * TMEM accumulation segment, 0 = whole K), "gemm_impl" (0 = tcgen05, 1 = CUDA-core checker),
 * "num_sms" (persistent grid size, 0 = all), "profile" (1 = record per-kernel CUDA events),
 * "mid_band" (thousandths of a pixel, default 26 = 2x the largest 1-pass/3-pass difference measured over 125k DISTINCT
 * coordinates, profiles/r02_band_stats.json; with mid_passes = 3 every row is first computed 1-pass and
 * only rows with a coordinate within the band of an integer -- where trunc(mid) could differ from the
 * reference -- are re-computed 3-pass; 0 = 3-pass for every row), "fuse_gather" (conv1 A operand of the 1-pass launches; default 3 = per-image window map + strided TMA boxes, no
 * producer warps; 1 = gathered in producer warps; 2 = first-generation fused kernel; 0 = separate gather kernel + TMA
 * of a materialised patch tensor; all four give bit-identical results except 0, which skips one fp16 rounding),
 * "nc_impl" (default 1: NeighConsensus on the tensor cores, nc_umma.cu; 0: fp32 CUDA-core kernels, B grid <= 3072 cells),
 * "nc_l2_mode" (layout of NC layer 2's block of hidden lines: 0 = chosen per shape, 1 = one haloed block per tile,
 * 2 = one block per column tap; bit-identical results), "unique_impl" (default 1: rank sort over the whole GPU for lists of
 * up to 8192 rows; 0: single-block bitonic network; identical results), "fc_impl" (default 1: the 512-512 and 512-256 Linear layers run on
 * the tensor cores, 3-pass; 0: fp32 CUDA-core FC kernel), "gemm_pair" (bitmask of GEMM launches that run
 * on the CTA-pair kernel -- tcgen05.mma.cta_group::2, one M=256 tile over the two SMs of a TPC, bit-identical
 * results: 1 = 1-pass convs, 2 = 3-pass convs, 4 = FC, 8 = correlation, 16 = p2p_test_gemm, 32 = fused-gather conv1;
 * default 35 = all conv launches). */
P2P_API int p2p_set_option(p2p_handle_t h, const char* key, int value);
P2P_API int p2p_get_option(p2p_handle_t h, const char* key, int* value);
/* Number of kernel launches enqueued by this handle since creation (bench.py's gpu_launches). */
P2P_API int p2p_launch_count(p2p_handle_t h, long long* count);

/* Per-kernel-group device times (CUDA events on the launching stream), accumulated while the
 * "profile" option is 1.  Synchronises on the recorded events, returns ms and launch-group counts
 * per kind, and clears the log. */
enum {
  P2P_PROF_L2NORM = 0, P2P_PROF_CORR = 1, P2P_PROF_MUTUAL = 2, P2P_PROF_NC = 3, P2P_PROF_PROPOSALS = 4,
  P2P_PROF_PREP = 5, P2P_PROF_GATHER_MID = 6, P2P_PROF_CONV1_MID = 7, P2P_PROF_CONV2_MID = 8, P2P_PROF_FC_MID = 9,
  P2P_PROF_GATHER_FINE = 10, P2P_PROF_CONV1_FINE = 11, P2P_PROF_CONV2_FINE = 12, P2P_PROF_FC_FINE = 13,
  P2P_PROF_GATHER_BAND = 14, P2P_PROF_CONV1_BAND = 15, P2P_PROF_CONV2_BAND = 16, P2P_PROF_FC_BAND = 17,
  P2P_PROF_FLAG = 18, P2P_PROF_KINDS = 19
};
P2P_API int p2p_profile_read(p2p_handle_t h, float* ms_by_kind, int* count_by_kind, int nkinds);

/* ---- coarse stage: Patch2Pix.forward_coarse_match (networks/patch2pix.py:120-136) -------------
 * feat1 [c,h1,w1], feat2 [c,h2,w2] fp32 (one batch item).  ksize 1 or 2.
 * corr4d_out [hp1*wp1, hp2*wp2] fp32 with hp = h/ksize (the final MutualMatching output).
 * delta_code_out (ksize 2 only, may be NULL for ksize 1): uint8 [hp1*wp1, hp2*wp2],
 * code = ((di*k+dj)*k+dk)*k+dl of maxpool4d (networks/modules.py:11-34).
 * Optional stage taps for parity tests (may be NULL): pooled_out (after maxpool4d),
 * ncn_out (after NeighConsensus, before the second MutualMatching). */
P2P_API int p2p_coarse(p2p_handle_t h, const float* feat1, const float* feat2, int c, int h1, int w1, int h2, int w2,
               int ksize, float* corr4d_out, uint8_t* delta_code_out, float* pooled_out, float* ncn_out,
               void* stream);

/* Same, for a channels-last fp16 layer-3 map: feat1 [h1][w1][c], feat2 [h2][w2][c] (what an fp16 / channels_last
 * backbone emits; the L2-normalise + K-major re-layout pass then needs no transpose).  Identical arithmetic from there on. */
P2P_API int p2p_coarse_nhwc16(p2p_handle_t h, const void* feat1_nhwc16, const void* feat2_nhwc16, int c, int h1, int w1, int h2,
               int w2, int ksize, float* corr4d_out, uint8_t* delta_code_out, float* pooled_out, float* ncn_out,
               void* stream);

/* maxpool4d's four int64 delta tensors (max_i,max_j,max_k,max_l) from the packed code. */
P2P_API int p2p_delta_unpack(p2p_handle_t h, const uint8_t* code, long long n, int ksize, int64_t* di, int64_t* dj,
                     int64_t* dk, int64_t* dl, void* stream);
/* Inverse, for callers that hand in reference-style delta4d tensors. */
P2P_API int p2p_delta_pack(p2p_handle_t h, const int64_t* di, const int64_t* dj, const int64_t* dk, const int64_t* dl,
                   long long n, int ksize, uint8_t* code, void* stream);

/* MutualMatching alone (networks/ncn/model.py:157-176) on [nA, nB]. */
P2P_API int p2p_mutual_matching(p2p_handle_t h, const float* in, int nA, int nB, float* out, void* stream);
/* NeighConsensus alone (networks/ncn/model.py:145-155) on [hA,wA,hB,wB]. */
P2P_API int p2p_neigh_consensus(p2p_handle_t h, const float* in, int hA, int wA, int hB, int wB, float* out, void* stream);

/* ---- proposals: Patch2Pix.cal_coarse_matches (networks/patch2pix.py:340-375, sort=False) over
 * corr_to_matches (networks/ncn/extract_ncmatches.py:6-94).  corr4d [hA*wA, hB*wB]; delta_code may be
 * NULL (ksize 1).  matches_out int64 [hB*wB + hA*wA, 4] rows (x1,y1,x2,y2): first the best A cell for
 * every B cell, then the best B cell for every A cell; scores_out fp32 same length. */
P2P_API int p2p_proposals(p2p_handle_t h, const float* corr4d, const uint8_t* delta_code, int hA, int wA, int hB, int wB,
                  int ksize, int upsample, int center, int do_softmax, int64_t* matches_out, float* scores_out,
                  void* stream);

/* ---- the np.unique part of filter_coarse (networks/utils.py:38-50): lexicographically sorted
 * first-occurrence indices of the distinct rows (mutual != 0: only rows occurring more than once).
 * rows int64 [n,4] with coordinates in [0,65535], n <= 4 Mi (lists beyond 16384 rows sort in global scratch); ids_out int32 [n]; count_out int32 [4] =
 * {number of ids, 1 if a coordinate was out of range, number of those ids whose score > thres,
 * number of ALL rows whose score > thres}; the last two (scores may be NULL -> 0) let the caller
 * evaluate filter_coarse's score threshold (utils.py:53) without a second device sync. */
P2P_API int p2p_unique_rows(p2p_handle_t h, const int64_t* rows, int n, int mutual, const float* scores, float thres,
                    int32_t* ids_out, int32_t* count_out, void* stream);

/* ---- the index arithmetic that follows np.unique in filter_coarse (networks/utils.py:51-69: matches[ids][ids2],
 * scores likewise) fused with Patch2Pix.shift_to_anchors (networks/patch2pix.py:377-402).
 * Output row r <- rows[ids[sel[r]]] (ids, sel int32 DEVICE, either may be NULL = identity); m output rows.
 * panc 8: anchors_out int64 [m*8,4] = every selected row + the reference's 8-row template
 * ((-p,-p,0,0),(p,-p,0,0),(-p,p,0,0),(p,p,0,0),(0,0,-p,-p),(0,0,p,-p),(0,0,-p,p),(0,0,p,p)); panc 1: anchors_out unused.
 * matches_out int64 [m,4] / scores_out fp32 [m] may be NULL. */
P2P_API int p2p_select_anchor(p2p_handle_t h, const int64_t* rows, const float* scores, const int32_t* ids,
                      const int32_t* sel, int m, int panc, int pshift, int64_t* matches_out, float* scores_out,
                      int64_t* anchors_out, void* stream);

/* ---- refine: Patch2Pix.forward_fine_match for one batch item (networks/patch2pix.py:157-218):
 * select_local_patch_feats + L2 normalise + FeatRegressNet + parse_regressor_out.
 * feats1/feats2: HOST arrays of 4 DEVICE pointers: image [3,H,W], conv1-relu [64,H/2,W/2],
 * layer1 [64,H/4,W/4], layer2 [128,H/8,W/8] (ResNet.forward_all levels 0..3, networks/resnet.py:138-157).
 * `p2p_refine_prepare` must be called once per image pair before p2p_refine (it builds channels-last
 * copies and norm maps); matches_in [n,4] int64 (is_float = 0) or fp32 (is_float = 1);
 * matches_out fp32 [n,4]; probs_out fp32 [n]. */
P2P_API int p2p_refine_prepare(p2p_handle_t h, const float* const* feats1, const float* const* feats2, int H1, int W1,
                       int H2, int W2, void* stream);
/* Same, with levels 1..3 as channels-last fp16 maps [h][w][C] (level 0, the image, stays [3,H,W] fp32). */
P2P_API int p2p_refine_prepare_nhwc16(p2p_handle_t h, const void* const* feats1, const void* const* feats2, int H1, int W1,
                       int H2, int W2, void* stream);
P2P_API int p2p_refine(p2p_handle_t h, int which, const void* matches_in, int is_float, int n, float* matches_out,
               float* probs_out, void* stream);

/* ---- tail of estimate_matches (utils/eval/model_helper.py:97-109): inlier filter `scores > io_thres` ("keep everything
 * if nothing passes"), row order preserved, and `upscale * matches` in float64, so that ONE device->host copy returns
 * the final result.  fine fp32 [n,4] (NULL for eval_type 'coarse': the refined columns repeat the coarse ones), scores
 * fp32 [n], coarse int64 [n,4], upscale4 HOST double[4] = (sx1, sy1, sx2, sy2).  packed_out DEVICE double [n*9 + 1]:
 * rows (x1,y1,x2,y2 refined, score, x1,y1,x2,y2 coarse), packed_out[n*9] = number of rows kept. */
P2P_API int p2p_finalize_matches(p2p_handle_t h, const float* fine, const float* scores, const int64_t* coarse, int n,
                         float io_thres, const double* upscale4, double* packed_out, void* stream);

/* ---- image preprocessing: the tensor half of load_im_flexible (utils/datasets/preprocess.py:32-60):
 * transforms.functional.resize(img, (ht, wt), Image.BICUBIC) -> ToTensor -> Normalize(ImageNet mean/std) for a decoded
 * 8-bit RGB image.  Pillow's 8-bit resampling (fixed-point, antialiased bicubic, horizontal then vertical pass) is
 * reproduced bit-exactly.  rgb_hwc: DEVICE uint8 [ho][wo][3]; out_chw: DEVICE fp32 [3][ht][wt];
 * resized_hwc_out (optional, may be NULL): DEVICE uint8 [ht][wt][3], the resized image before normalisation.
 * (ht, wt) come from cal_rescale_size (preprocess.py:83-91), computed by the caller. */
P2P_API int p2p_preprocess_image(p2p_handle_t h, const uint8_t* rgb_hwc, int ho, int wo, int ht, int wt, float* out_chw,
                         uint8_t* resized_hwc_out, void* stream);

/* ---- bring-up / accuracy probe: C[M,N] = alpha * A[M,K] B[N,K]^T on the tcgen05 path with the
 * same operand format as the hot path (fp32 inputs are split to fp16 hi/lo on the device).
 * a, b, c are DEVICE fp32; K % 64 == 0. */
P2P_API int p2p_test_gemm(p2p_handle_t h, const float* a, const float* b, float* c, int M, int N, int K, int passes,
                  int seg_len, float in_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* P2P_B200_H_ */
