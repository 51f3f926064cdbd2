// Internal launcher declarations shared by the translation units of libp2p_b200.so.
#pragma once
#include "common.cuh"

namespace p2p {

// ---- coarse.cu ---------------------------------------------------------------------------------
int launch_l2norm_perm(const float* in, float* out, int C, int h, int w, int ksize, cudaStream_t st);
// K-major fp16 hi/lo variant for the tensor-core correlation: out[q][c] = split(kActScale * f/|f|)
int launch_l2norm_perm_kmajor_pair(const float* in1, const float* in2, __half* hi1, __half* lo1, __half* hi2, __half* lo2,
                                   int C, int h1, int w1, int h2, int w2, int ksize, cudaStream_t st);
int launch_l2norm_perm_kmajor_pair_nhwc16(const __half* in1, const __half* in2, __half* hi1, __half* lo1, __half* hi2, __half* lo2,
                                          int C, int h1, int w1, int h2, int w2, int ksize, cudaStream_t st);
int launch_split_rows(const float* in, __half* hi, __half* lo, size_t n, float scale, cudaStream_t st);
int launch_delta_pack(const long long* di, const long long* dj, const long long* dk, const long long* dl, size_t n,
                      int ks, uint8_t* code, cudaStream_t st);
int launch_corr_pool_simt(const float* fa, const float* fb, int C, int n1, int n2, int ksize, float* out,
                          uint8_t* code, cudaStream_t st);
int launch_delta_unpack(const uint8_t* code, size_t n, int ks, long long* di, long long* dj, long long* dk,
                        long long* dl, cudaStream_t st);
// absmax (optional): device word that receives the float bits of max |out| (activation scale of the tensor-core NC)
int launch_mutual_matching(const float* x, int nA, int nB, float* rowmax, unsigned int* colmax, float* out,
                           unsigned int* absmax, cudaStream_t st);
// second half only: rowmax / colmax were produced by another kernel (nc_combine_kernel)
int launch_mutual_apply(const float* x, int nA, int nB, const float* rowmax, const unsigned int* colmax, float* out,
                        unsigned int* absmax, cudaStream_t st);
int launch_neigh_consensus(const float* x, int hA, int wA, int hB, int wB, const float* w1p, const float* b1p,
                           const float* w2p, float b2, float* hidden, float* out, cudaStream_t st);
int launch_proposals(const float* corr, const uint8_t* code, int hA, int wA, int hB, int wB, int ksize, int upsample,
                     int center, int do_softmax, long long* matches, float* scores, cudaStream_t st);
size_t unique_rows_scratch_bytes(int n);
size_t unique_rank_scratch_bytes();     // zero-initialised, handle-owned scratch of the rank-sort path (n <= 8192)
int launch_unique_rows(const long long* rows, int n, int mutual, const float* scores, float thres, int* ids_out,
                       int* count_out, unsigned char* gscratch, int* rank_scratch, cudaStream_t st);
int launch_select_anchor(const long long* rows, const float* scores, const int* ids, const int* sel, int m, int panc,
                         int pshift, long long* matches_out, float* scores_out, long long* anchors_out, cudaStream_t st);

// ---- nc_umma.cu: NeighConsensus on the tensor cores -------------------------------------------------------
struct NcUmmaWeights {
  char* blob = nullptr;           // one allocation backing the two operand images
  __half* img1 = nullptr;         // layer 1 weights, laid out exactly as in shared memory
  __half* img2 = nullptr;         // layer 2 weights
  float wsum1 = 0.f, b1max = 0.f; // bound of the hidden activations: b1max + wsum1 * max|x|
  float inv_sw1 = 1.f, inv_sw2 = 1.f;
};
int nc_umma_pack(const float* w1p, const float* b1p, const float* w2p, NcUmmaWeights& W);
size_t nc_umma_scratch_bytes(size_t V);
size_t nc_umma_xp_bytes(int hA, int wA, int hB, int wB);
int launch_absmax(const float* x, size_t n, unsigned int* out, cudaStream_t st);
int launch_neigh_consensus_umma(const float* x, int hA, int wA, int hB, int wB, const NcUmmaWeights& W, const float* b1p,
                                float b2, const unsigned int* xmax, uint32_t* xp, __half* hidden, float* partial,
                                float* out, float* rowmax, unsigned int* colmax, int l2_mode, int num_sms, cudaStream_t st);

// ---- preprocess.cu: PIL-exact bicubic resize + ToTensor + Normalize (utils/datasets/preprocess.py:32-60) ----
struct PreprocessCoefs {
  int ho = 0, wo = 0, ht = 0, wt = 0, ksx = 0, ksy = 0;
  size_t o_bx = 0, o_kx = 0, o_by = 0, o_ky = 0;
  int* d = nullptr;               // device: [bounds x][coefs x][bounds y][coefs y]
};
int preprocess_build_coefs(int ho, int wo, int ht, int wt, PreprocessCoefs& C);
int launch_preprocess(const uint8_t* rgb, const PreprocessCoefs& C, const float mean[3], const float stdv[3], float* out,
                      uint8_t* resized_u8, uint8_t* tmp, cudaStream_t st);

// ---- refine.cu ---------------------------------------------------------------------------------
// Activation scale applied before the fp16 hi/lo split of the L2-normalised patch features.
constexpr float kActScale = 4096.f;
constexpr int kPatchPos = 256;    // 4 parity planes x 8 x 8 window positions
constexpr int kMainCh = 512;      // (64 + 64 + 128) x 2 images
constexpr int kRgbK = 64;         // 9 taps x 2 images x 3 channels = 54, padded to 64
constexpr int kConv1Steps = 73;   // 9 taps x 8 chunks + 1 rgb chunk (K = 73*64 = 4672)
constexpr int kConv2Steps = 72;   // 9 taps x 8 chunks          (K = 4608)

struct PairFeatures {             // per image: channels-last copies + squared-norm maps
  const float* img;               // [3][H][W]   (level 0 stays NCHW)
  float* nhwc[3];                 // levels 1..3: [h][w][C], C = 64, 64, 128
  __half* nhwc16[3];              // same, fp16, each pixel divided by its own level norm sqrt(nsq[l+1])
  float* nsq[4];                  // levels 0..3: [h][w]
  int H, W;
  // full-resolution window map (fuse_gather = 3): every pixel's patch-normalised 256-channel vector, replicate-padded by
  // kMapPad pixels, [H + 2 pad][W + 2 pad][256] fp16; rgbn: the normalised rgb triple [..][4] fp16
  __half* wmap;
  __half* rgbn;
};
constexpr int kMapPad = 16;
int launch_window_map(const PairFeatures pf[2], cudaStream_t st);

int launch_feature_prep_pair(const float* const feats1[4], const float* const feats2[4], const int H[2], const int W[2],
                             PairFeatures out[2], int fmt, cudaStream_t st);
// rowmap/d_count (optional, device): process only rows rowmap[0..*d_count) (patch slot b <- row rowmap[b]).
int launch_patch_gather(const PairFeatures& f1, const PairFeatures& f2, const void* matches, int is_float, int N,
                        __half* p_hi, __half* p_lo, __half* rgb_hi, __half* rgb_lo, const int* rowmap,
                        const int* d_count, cudaStream_t st);
// Rows whose refined coordinates sit within `tau` px of an integer (and whose offset is not the exact
// relu-clamped -8) are collected, in ascending order, into rowmap / d_count.
int launch_flag_risky(const void* matches_in, int is_float, const float* raw, int N, float tau, float eps_o, int W1,
                      int H1, int W2, int H2, int* rowmap, int* d_count, unsigned long long* totals, cudaStream_t st);

struct FcWeights {                // BN folded, transposed to [in][out] for coalesced reads
  float *w1t, *b1, *w2t, *b2, *w3t, *b3;
};
int launch_fc_parse(const float* pooled, const FcWeights& fc, const void* matches_in, int is_float, int N, int W1,
                    int H1, int W2, int H2, float* matches_out, float* probs_out, float* raw_out, const int* rowmap,
                    const int* d_count, cudaStream_t st);

int launch_finalize_matches(const float* fine, const float* scores, const long long* coarse, int N, float io_thres,
                            const double up[4], double* packed, cudaStream_t st);

// Tensor-core FC path helpers: pooled fp32 -> fp16 hi/lo A operand; final Linear(256,5) + parse_regressor_out.
constexpr float kFcActScale = 16.f;
int launch_pooled_split(const float* pooled, int n, __half* hi, __half* lo, const int* d_count, cudaStream_t st);
int launch_fc3_parse(const __half* h2_hi, const __half* h2_lo, const float* w3t, const float* b3, const void* matches_in,
                     int is_float, int N, int W1, int H1, int W2, int H2, float* matches_out, float* probs_out,
                     float* raw_out, const int* rowmap, const int* d_count, cudaStream_t st);

// One k-step of an implicit GEMM: where the [128 rows x 64 ch] A box starts and which K offset of
// the K-major weight matrix it multiplies.
struct KStep {
  short c0;                       // channel coordinate of the A box
  signed char x, y;               // spatial start (may be -1: TMA zero-fills out-of-bounds)
  signed char plane;              // parity plane (conv1) or 0
  signed char kind;               // 0 = main activation tensor, 1 = rgb im2col tensor
  short pad;
  int bk;                         // K coordinate in the weight matrix
};

// CUDA-core debug GEMM over exactly the tensors the tcgen05 kernel consumes (bring-up checker;
// not a product path: selected only by p2p_set_option("gemm_impl", 1)).
struct GemmOperands {
  const __half *a_hi, *a_lo;      // main A tensor [N][planes][8][8][512]
  const __half *r_hi, *r_lo;      // rgb A tensor  [N][64][64]  (nullptr for conv2)
  const __half *b_hi, *b_lo;      // weights [512][Ktot], K-major
  int planes;                     // 4 (conv1) or 1 (conv2)
  int ktot;                       // K extent of B
  int n_patches;
  int passes;                     // 1: hi*hi ; 3: hi*hi + lo*hi + hi*lo
  const KStep* steps;             // device array
  int nsteps;
};
struct ConvEpilogue {
  const float* scale;             // [512] = 1 / (act_scale * w_scale[o])
  const float* bias;              // [512]
  int mode;                       // 0: write y1 hi/lo (scaled by y_scale); 1: relu + 8x8 max -> pooled
  float y_scale;
  __half *y_hi, *y_lo;            // [N][8][8][512]
  float* pooled;                  // [N][512]
};
int launch_conv_gemm_simt(const GemmOperands& g, const ConvEpilogue& e, cudaStream_t st);

}  // namespace p2p
