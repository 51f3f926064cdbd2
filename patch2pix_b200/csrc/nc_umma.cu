// NeighConsensus (symmetric 2-layer Conv4d 1 -> 16 -> 1, k = 3, ReLU after each layer) on the tcgen05 tensor cores.
//
// Reference semantics (file:line relative to the reference repo):
//   NeighConsensus.forward   networks/ncn/model.py:145-155   conv(x) + conv(x^T)^T, shared weights
//   Conv4d / conv4d          networks/ncn/conv4d.py:12-130   true 4D cross-correlation, zero "same" padding
//
// conv(x) + conv(x^T)^T equals two independent nets on the SAME input, the second with tap axes (a,b) <-> (d,e)
// swapped (api.cu packs both: w1p / w2p [81 taps][32 = 16 ch of net 0 | 16 ch of net 1]).
//
// Both layers are skinny GEMMs whose A operand is an im2col matrix that producer warps build directly in shared
// memory in the 128B-swizzled K-major layout tcgen05.mma consumes; nothing but x, the hidden tensor and the
// partial maps touches HBM and no FMA-pipe inner loop is left:
//
//   layer 1   rows = 128 consecutive 4D cells, K = 81 taps (padded to 128), N = 32 channels (both nets).
//             Epilogue: + bias, ReLU, re-scale, fp16 hi/lo split -> hidden[cell][64 fp16] =
//             [net0 hi 16 | net0 lo 16 | net1 hi 16 | net1 lo 16] (one 128-byte line per cell).
//   layer 2   16 -> 1 channels would be an N = 1 GEMM.  Instead, per hidden cell a' and per net, the 9 PARTIAL maps
//             P_(ta,tb)[a'][b] = sum over the 9 B-taps and 16 channels are one GEMM with K = 9 x 16 = 144, N = 9
//             (padded to 16); one operand atom per B-tap whose rows are verbatim copies of 128-byte hidden lines.
//   combine   out[a][b] = sum_net relu(b2 + sum_(ta,tb) P_(ta,tb)[a + (ta-1, tb-1)][b])  (fixed summation order:
//             deterministic), fused with the row/column maxima of the MutualMatching that follows.
//
// Precision: fp16 hi/lo operand pairs, three MMAs per product (lo*hi + hi*lo + hi*hi, fp32 accumulate in TMEM):
// products good to ~2^-22, i.e. fp32-grade (chains are 18 / 27 MMAs long, so the accumulator's round-toward-zero
// stays below 1e-6 relative).  Activations are scaled by powers of two derived ON THE DEVICE from max|x| (and from
// a weight-norm bound for the hidden tensor), so any input range is safe in fp16.
//
// Warp-specialised persistent kernels (MMA issuer, TMEM allocator, 4 epilogue warps = TMEM lane quadrants, producer
// warps), mbarrier ring of operand stages, two TMEM accumulator slots so that the epilogue of tile i overlaps the MMAs
// of tile i+1.
#include <math.h>

#include <type_traits>
#include <vector>

#include "kernels.h"
#include "umma_ptx.cuh"

namespace p2p {

constexpr int kNcAtom = 128 * 128;   // bytes of one [128 rows x 64 fp16] swizzled operand atom

struct NcParams {
  int hA, wA, hB, wB, nA, nB;
  long long V;                 // nA * nB cells
  const float* x;              // [V] input (after the first MutualMatching)
  const unsigned int* xmax;    // device: float bits of max |x|
  __half* hidden;              // [V][64]
  float* partial;              // [2 nets][9][V]
  const __half* wimg;          // weight operand image, laid out exactly as in shared memory
  const float* b1p;            // [32]
  float wsum1, b1max;          // max_c sum_taps |w1|, max |b1|: bound of the hidden activations
  float inv_sw1, inv_sw2;      // 1 / (power-of-two weight scales)
  int tiles;
  int l1_bufs;                 // layer-1 staging buffers: 2 (copies of tile i+1 overlap tile i) or 1 (wide B grids)
};

// Power-of-two activation scales: max|x| * sx and (hidden bound) * sh land in [2048, 4096).
__device__ __forceinline__ void nc_scales(const NcParams& p, float& sx, float& sh) {
  float xmax = __uint_as_float(__ldg(p.xmax));
  if (!(xmax > 0.f) || !isfinite(xmax)) xmax = 1.f;
  int e;
  frexpf(xmax, &e);
  sx = ldexpf(1.f, min(12 - e, 60));
  float hb = fmaf(p.wsum1, xmax, p.b1max);
  if (!(hb > 0.f) || !isfinite(hb)) hb = 1.f;
  frexpf(hb, &e);
  sh = ldexpf(1.f, min(12 - e, 60));
}

__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = fminf(fmaxf(v[2 * i], -65504.f), 65504.f), b = fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f);
    h[i] = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h[i]);
    l[i] = __floats2half2_rn(a - f.x, b - f.y);
  }
  hi = *reinterpret_cast<uint4*>(h);
  lo = *reinterpret_cast<uint4*>(l);
}

// operands that are bounded by construction (|x * sx| < 4096): no saturation needed
__device__ __forceinline__ void split8_bounded(const float* v, uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 f = __half22float2(h[i]);
    l[i] = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
  }
  hi = *reinterpret_cast<uint4*>(h);
  lo = *reinterpret_cast<uint4*>(l);
}

// ------------------------------------------------------------------------------------------------
// layer 1.  Tile = (A cell a, 128 consecutive B cells).  The 9 A-neighbours' B rows touched by the tile are staged in
// shared memory with cp.async as a zero-PADDED 2-D block (one extra row above/below, one extra column left/right,
// zero outside the volume; double-buffered: tile i+1 is in flight while tile i is built), so a tap is a plain
// `ld.shared [row base + tk*pitch + tl]` with no validity logic.  Sixteen producer warps (four threads per tile row,
// each a quarter of the 11 tap chunks) scale, split to fp16 hi/lo and store the swizzled operand chunks.
// The three products of the hi/lo split (lo*hi, hi*lo, hi*hi) accumulate in three SEPARATE TMEM blocks -- three
// independent MMA chains instead of one 18-deep dependent chain of tiny (N = 32) MMAs -- and are summed by the epilogue.
// 768 threads (warp 1 MMA, 2 TMEM, 4..7 epilogue, 8..23 producers), 1 CTA per SM, 4 x 32 KB operand stages.
// ------------------------------------------------------------------------------------------------
constexpr int kL1Stages = 4;
constexpr int kL1Threads = 768, kL1Producers = 512;

__host__ __device__ inline int nc_l1_rows(int wB) { return 127 / wB + 4; }      // B rows a tile can touch, + halo

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(kL1Threads, 1) nc_l1_umma_kernel(const __grid_constant__ NcParams p) {
  constexpr int STAGE_BYTES = 2 * kNcAtom;         // A_hi + A_lo of one atom
  constexpr int WATOM = 32 * 128;
  constexpr uint32_t IDESC = make_idesc_f16(128, 32);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem + kL1Stages * STAGE_BYTES;                          // [hi|lo][atom] weight images, 16 KB
  float* xs = reinterpret_cast<float*>(wsm + 4 * WATOM);                  // [2 buffers][9][rows][pitch]
  __shared__ __align__(8) uint64_t full_bar[kL1Stages];
  __shared__ __align__(8) uint64_t empty_bar[kL1Stages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.tiles;
  const int TB = (p.nB + 127) >> 7;
  const int PW = p.wB + 2, NR = nc_l1_rows(p.wB), DS = NR * PW;           // pitch, rows, floats per A-neighbour block

  for (int i = threadIdx.x; i < kL1Stages * STAGE_BYTES / 16; i += kL1Threads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < 4 * WATOM / 16; i += kL1Threads)
    reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(p.wimg) + i);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kL1Stages; ++i) {
      mbar_init(&full_bar[i], kL1Producers);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 1) {
    if (lane == 0) {
      int it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tl) {
        const int slot = tl & 1;
        mbar_wait(&tempty_bar[slot], ((uint32_t)(tl >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(slot * 96);            // three 32-column blocks: lo*hi, hi*lo, hi*hi
#pragma unroll
        for (int atom = 0; atom < 2; ++atom, ++it) {
          const int s = it % kL1Stages;
          mbar_wait(&full_bar[s], (uint32_t)(it / kL1Stages) & 1u);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint64_t a_hi = make_sw128_desc(sa), a_lo = make_sw128_desc(sa + kNcAtom);
          const uint32_t wb = smem_u32(wsm) + (uint32_t)(atom * WATOM);
          const uint64_t w_hi = make_sw128_desc(wb), w_lo = make_sw128_desc(wb + 2 * WATOM);
          const int nk = atom == 0 ? 4 : 2;           // taps 64..80 live in the first two K16 slices of atom 1
          for (int kk = 0; kk < nk; ++kk) {
            const uint32_t acc = (atom > 0 || kk > 0) ? 1u : 0u;
            umma_f16(d0, a_lo + 2 * kk, w_hi + 2 * kk, IDESC, acc);
            umma_f16(d0 + 32, a_hi + 2 * kk, w_lo + 2 * kk, IDESC, acc);
            umma_f16(d0 + 64, a_hi + 2 * kk, w_hi + 2 * kk, IDESC, acc);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[slot]);
      }
    }
  } else if (warp >= 8) {
    // ===================== producers: 512 threads = 128 tile rows x 4 chunk quarters =====================
    const int ptid = threadIdx.x - 256;
    const int r = ptid & 127;
    const int qt = ptid >> 7;          // warp-uniform: chunks qt and qt + 4 of atom 0, chunk qt of atom 1 (qt < 3)
    float sx, sh;
    nc_scales(p, sx, sh);
    auto issue = [&](int tile, int buf) {
      const int a = tile / TB, b0 = (tile - a * TB) << 7;
      const int ia = a / p.wA, ja = a - ia * p.wA;
      const int kr0 = b0 / p.wB - 1;                          // B row of block row 0
      float* dst = xs + (size_t)buf * 9 * DS;
      for (int idx = ptid; idx < 9 * DS; idx += kL1Producers) {
        const int d = idx / DS, rem = idx - d * DS;
        const int rr = rem / PW, cc = rem - rr * PW;
        const int i2 = ia + d / 3 - 1, j2 = ja + d % 3 - 1;
        const int k = kr0 + rr, l = cc - 1;
        const bool ok = i2 >= 0 && i2 < p.hA && j2 >= 0 && j2 < p.wA && k >= 0 && k < p.hB && l >= 0 && l < p.wB;
        const float* src = ok ? p.x + (size_t)(i2 * p.wA + j2) * p.nB + k * p.wB + l : p.x;
        const unsigned sz = ok ? 4u : 0u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst + idx)), "l"(src), "r"(sz) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int it = 0, buf = 0;
    const bool dbuf = p.l1_bufs == 2;
    if (dbuf && (int)blockIdx.x < tiles) issue(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int next = tile + gridDim.x;
      if (!dbuf) {
        issue(tile, 0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else if (next < tiles) {
        issue(next, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");          // every producer's copies of this tile have landed
      const int a = tile / TB, b0 = (tile - a * TB) << 7;
      const int b = b0 + r;
      const bool rv = b < p.nB;
      const int k = rv ? b / p.wB : b0 / p.wB, l = rv ? b - k * p.wB : 0;
      const float sxr = rv ? sx : 0.f;                         // rows past the end of the B grid produce zeros
      // shared address of block element (row k - 1, column l - 1) of A-neighbour 0: tap (d, tk, tl) is at
      // + d * DS + tk * PW + tl floats
      const uint32_t base = smem_u32(xs + (size_t)buf * 9 * DS) + (uint32_t)(((k - (b0 / p.wB - 1) - 1) * PW + l) * 4);
      const uint32_t otk1 = (uint32_t)(PW * 4), otk2 = (uint32_t)(2 * PW * 4), ods = (uint32_t)(DS * 4);
      auto chunk = [&](auto ATOM, auto CH, uint8_t* st) {
        constexpr int atom = decltype(ATOM)::value, c = decltype(CH)::value;
        float val[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int t = (atom * 8 + c) * 8 + i;                  // compile-time
          float f = 0.f;
          if (t < 81) {
            const int d = t / 9, tk = (t / 3) % 3, tl = t % 3;
            f = lds_f32(base + (uint32_t)d * ods + (tk == 0 ? 0u : (tk == 1 ? otk1 : otk2)) + (uint32_t)(tl * 4)) * sxr;
          }
          val[i] = f;
        }
        uint4 hi, lo;
        split8_bounded(val, hi, lo);
        *reinterpret_cast<uint4*>(st + ((c ^ (r & 7)) << 4)) = hi;
        *reinterpret_cast<uint4*>(st + kNcAtom + ((c ^ (r & 7)) << 4)) = lo;
      };
      auto build = [&](auto QT) {
        constexpr int Q = decltype(QT)::value;
        {   // atom 0: chunks Q and Q + 4
          const int s = it % kL1Stages;
          mbar_wait(&empty_bar[s], ((uint32_t)(it / kL1Stages) & 1u) ^ 1u);
          uint8_t* st = smem + (size_t)s * STAGE_BYTES + r * 128;
          chunk(std::integral_constant<int, 0>{}, std::integral_constant<int, Q>{}, st);
          chunk(std::integral_constant<int, 0>{}, std::integral_constant<int, Q + 4>{}, st);
          fence_proxy_async();
          mbar_arrive(&full_bar[s]);
          ++it;
        }
        {   // atom 1: chunk Q (taps 64 + 8Q ..; Q = 3 has nothing to write, chunks 3..7 stay zero)
          const int s = it % kL1Stages;
          mbar_wait(&empty_bar[s], ((uint32_t)(it / kL1Stages) & 1u) ^ 1u);
          if (Q < 3) {
            uint8_t* st = smem + (size_t)s * STAGE_BYTES + r * 128;
            chunk(std::integral_constant<int, 1>{}, std::integral_constant<int, (Q < 3 ? Q : 0)>{}, st);
            fence_proxy_async();
          }
          mbar_arrive(&full_bar[s]);
          ++it;
        }
      };
      if (qt == 0) build(std::integral_constant<int, 0>{});
      else if (qt == 1) build(std::integral_constant<int, 1>{});
      else if (qt == 2) build(std::integral_constant<int, 2>{});
      else build(std::integral_constant<int, 3>{});
      asm volatile("bar.sync 1, 512;" ::: "memory");          // all reads of this buffer done before it is refilled
      if (dbuf) buf ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float sx, sh;
    nc_scales(p, sx, sh);
    const float inv = p.inv_sw1 / sx;
    int tl = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tl) {
      const int slot = tl & 1;
      mbar_wait(&tfull_bar[slot], (uint32_t)(tl >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 96);
      float acc[32], t1[32];
      tmem_ld32(taddr, acc);                       // lo*hi
      tmem_ld32(taddr + 32, t1);                   // hi*lo
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] += t1[c];
      tmem_ld32(taddr + 64, t1);                   // hi*hi
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] += t1[c];
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[slot]);     // accumulators are in registers: the slot can be refilled
      const int a = tile / TB, b = ((tile - a * TB) << 7) + row;
      if (b < p.nB) {
        const long long v = (long long)a * p.nB + b;
        uint4 o[8];
#pragma unroll
        for (int g = 0; g < 4; ++g) {                    // 8 channels at a time: net 0 ch 0-7, 8-15, net 1 ch 0-7, 8-15
          float hval[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) hval[c] = fmaxf(fmaf(acc[g * 8 + c], inv, __ldg(p.b1p + g * 8 + c)), 0.f) * sh;
          split8_bounded(hval, o[(g >> 1) * 4 + (g & 1)], o[(g >> 1) * 4 + 2 + (g & 1)]);   // [net][hi0 hi1 lo0 lo1]
        }
        uint4* dst = reinterpret_cast<uint4*>(p.hidden + v * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = o[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------
// layer 2.  Tile = 128 consecutive 4D cells, both nets.  One operand atom per B-TAP: the atom row of a cell is the
// complete 128-byte hidden line of its tap neighbour, K = [net][hi|lo][16 ch] -- so the producers are pure line copies
// (8 lanes per line: fully coalesced loads, conflict-free swizzled stores) and the (net, hi|lo) factor of an MMA is
// selected by the K16 slice of the descriptors: per tap and net  lo*hi + hi*lo + hi*hi  = 3 MMAs (M128 N16 K16).
// 512 threads (warp 1 MMA, 2 TMEM, 4..7 epilogue, 8..15 producers); ring of 8 x 16 KB stages.
// ------------------------------------------------------------------------------------------------
constexpr int kL2Stages = 8;

__global__ void __launch_bounds__(512, 1) nc_l2_umma_kernel(const __grid_constant__ NcParams p) {
  constexpr int WATOM = 16 * 128;
  constexpr uint32_t IDESC = make_idesc_f16(128, 16);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem + kL2Stages * kNcAtom;       // [9 taps][16 rows][K = net | hi,lo | ch] weight images, 18 KB
  __shared__ __align__(8) uint64_t full_bar[kL2Stages];
  __shared__ __align__(8) uint64_t empty_bar[kL2Stages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.tiles;

  for (int i = threadIdx.x; i < 9 * WATOM / 16; i += 512)
    reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(p.wimg) + i);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kL2Stages; ++i) {
      mbar_init(&full_bar[i], 8);                  // one arrival per producer warp
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 1) {
    if (lane == 0) {
      int it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tl) {
        const int slot = tl & 1;
        mbar_wait(&tempty_bar[slot], ((uint32_t)(tl >> 1) & 1u) ^ 1u);
        tc_fence_after();
        for (int t = 0; t < 9; ++t, ++it) {
          const int s = it % kL2Stages;
          mbar_wait(&full_bar[s], (uint32_t)(it / kL2Stages) & 1u);
          tc_fence_after();
          const uint64_t a = make_sw128_desc(smem_u32(smem + (size_t)s * kNcAtom));
          const uint64_t w = make_sw128_desc(smem_u32(wsm) + (uint32_t)(t * WATOM));
#pragma unroll
          for (int net = 0; net < 2; ++net) {
            // six independent accumulator chains per tile (net x product), 16 columns each, summed by the epilogue:
            // a single chain of 27 dependent N = 16 MMAs is bound by the MMA latency, not by the tensor pipe
            const uint32_t d_tmem = tmem_base + (uint32_t)(slot * 96 + net * 48);
            const uint64_t a_hi = a + 2 * (net * 2), a_lo = a + 2 * (net * 2 + 1);      // K16 slices of the line
            const uint64_t w_hi = w + 2 * (net * 2), w_lo = w + 2 * (net * 2 + 1);
            const uint32_t acc = t > 0 ? 1u : 0u;
            umma_f16(d_tmem, a_lo, w_hi, IDESC, acc);
            umma_f16(d_tmem + 16, a_hi, w_lo, IDESC, acc);
            umma_f16(d_tmem + 32, a_hi, w_hi, IDESC, acc);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[slot]);
      }
    }
  } else if (warp >= 8) {
    // ===================== producers: 256 threads = 32 rows x 8 chunks per pass, 4 passes per tap =====================
    const int ptid = threadIdx.x - 256;
    const int c = ptid & 7, r0 = ptid >> 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      unsigned mask[4];
      const uint4* base[4];
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const long long v = (long long)tile * 128 + ps * 32 + r0;
        const bool rv = v < p.V;
        const int b = rv ? (int)(v % p.nB) : 0;
        const int k = b / p.wB, l = b - k * p.wB;
        mask[ps] = rv ? ((k > 0 ? 1u : 0u) | 2u | (k + 1 < p.hB ? 4u : 0u) | (l > 0 ? 8u : 0u) | 16u | (l + 1 < p.wB ? 32u : 0u)) : 0u;
        base[ps] = reinterpret_cast<const uint4*>(p.hidden + (rv ? v : 0) * 64) + c;
      }
      auto load_tap = [&](int t, uint4* q) {
        const int tk = t / 3, tl = t - tk * 3;
        const unsigned need = (1u << tk) | (8u << tl);
        const long long off = ((long long)(tk - 1) * p.wB + (tl - 1)) * 8;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) q[ps] = ((mask[ps] & need) == need) ? __ldg(base[ps] + off) : make_uint4(0, 0, 0, 0);
      };
      // three taps' lines in flight per thread (the loads are latency-bound: 8 of 9 lines hit L1, one comes from L2/HBM)
      uint4 q[3][4];
      load_tap(0, q[0]);
      load_tap(1, q[1]);
      load_tap(2, q[2]);
#pragma unroll
      for (int t = 0; t < 9; ++t, ++it) {
        const int s = it % kL2Stages;
        mbar_wait(&empty_bar[s], ((uint32_t)(it / kL2Stages) & 1u) ^ 1u);
        uint8_t* st = smem + (size_t)s * kNcAtom;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const int row = ps * 32 + r0;
          *reinterpret_cast<uint4*>(st + row * 128 + ((c ^ (row & 7)) << 4)) = q[t % 3][ps];
        }
        if (t + 3 < 9) load_tap(t + 3, q[t % 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float sx, sh;
    nc_scales(p, sx, sh);
    const float inv = p.inv_sw2 / sh;
    int tl = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tl) {
      const int slot = tl & 1;
      mbar_wait(&tfull_bar[slot], (uint32_t)(tl >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 96);
      float acc[2][16];
#pragma unroll
      for (int net = 0; net < 2; ++net) {
        float t1[16];
        tmem_ld16(taddr + net * 48, acc[net]);               // lo*hi
        tmem_ld16(taddr + net * 48 + 16, t1);                // hi*lo
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[net][d] += t1[d];
        tmem_ld16(taddr + net * 48 + 32, t1);                // hi*hi
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[net][d] += t1[d];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[slot]);
      const long long v = (long long)tile * 128 + row;
      if (v < p.V) {
#pragma unroll
        for (int net = 0; net < 2; ++net)
#pragma unroll
          for (int d = 0; d < 9; ++d) p.partial[(size_t)(net * 9 + d) * p.V + v] = acc[net][d] * inv;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------
// max |x| (device-side activation scale of the tensor-core NC path)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, size_t n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order like their bits
}

int launch_absmax(const float* x, size_t n, unsigned int* out, cudaStream_t st) {
  P2P_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(unsigned int), st));
  const int blocks = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  absmax_kernel<<<blocks, 256, 0, st>>>(x, n, out);
  P2P_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// combine: out[a][b] = sum_net relu(b2 + sum_(ta,tb) P[net][ta*3+tb][a + (ta-1, tb-1)][b]), fused with the
// row / column maxima of the MutualMatching that follows (rowmax[a] = max_b, colmax[b] = max_a).
// Block = kCombineRows A cells; the 9 neighbour offsets are block-uniform.
// ------------------------------------------------------------------------------------------------
constexpr int kCombineRows = 2;      // A cells per block: nA / 2 blocks keep every SM busy with several blocks

__global__ void __launch_bounds__(256) nc_combine_kernel(const float* __restrict__ P, int hA, int wA, int nB, float b2,
                                                        float* __restrict__ out, float* __restrict__ rowmax,
                                                        unsigned int* __restrict__ colmax) {
  constexpr int R = kCombineRows;
  __shared__ float red[8][R];
  const int nA = hA * wA;
  const size_t V = (size_t)nA * nB;
  const int r0 = blockIdx.x * R;
  float rm[R];
#pragma unroll
  for (int r = 0; r < R; ++r) rm[r] = -INFINITY;
  for (int col = threadIdx.x; col < nB; col += 256) {
    float cm = -INFINITY;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int a = r0 + r;
      if (a >= nA) continue;
      const int ia = a / wA, ja = a - ia * wA;
      float tot = 0.f;
#pragma unroll
      for (int net = 0; net < 2; ++net) {
        float acc = b2;
#pragma unroll
        for (int d = 0; d < 9; ++d) {
          const int i2 = ia + d / 3 - 1, j2 = ja + d % 3 - 1;
          if (i2 >= 0 && i2 < hA && j2 >= 0 && j2 < wA)
            acc += __ldg(P + (size_t)(net * 9 + d) * V + (size_t)(i2 * wA + j2) * nB + col);
        }
        tot += fmaxf(acc, 0.f);
      }
      out[(size_t)a * nB + col] = tot;
      rm[r] = fmaxf(rm[r], tot);
      cm = fmaxf(cm, tot);
    }
    if (colmax != nullptr) atomicMax(colmax + col, f2ord(cm));
  }
  if (rowmax == nullptr) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float v = warp_max(rm[r]);
    if (lane == 0) red[wid][r] = v;
  }
  __syncthreads();
  if (threadIdx.x < R && r0 + threadIdx.x < nA) {
    float v = red[0][threadIdx.x];
#pragma unroll
    for (int wv = 1; wv < 8; ++wv) v = fmaxf(v, red[wv][threadIdx.x]);
    rowmax[r0 + threadIdx.x] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static float pow2_scale_for(float maxabs, float target_hi) {   // power of two s: maxabs * s in [target_hi/2, target_hi)
  if (!(maxabs > 0.f) || !isfinite(maxabs)) return 1.f;
  int e, et;
  frexpf(maxabs, &e);
  frexpf(target_hi, &et);
  return ldexpf(1.f, et - 1 - e);
}

// element (n, k) of a [rows][64] K-major 128B-swizzled atom
static inline size_t sw128_index(int n, int k) { return (size_t)n * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7)); }

int nc_umma_pack(const float* w1p, const float* b1p, const float* w2p, NcUmmaWeights& W) {
  float m1 = 0.f, m2 = 0.f, wsum = 0.f, b1max = 0.f;
  for (int i = 0; i < 81 * 32; ++i) {
    m1 = fmaxf(m1, fabsf(w1p[i]));
    m2 = fmaxf(m2, fabsf(w2p[i]));
  }
  for (int c = 0; c < 32; ++c) {
    float s = 0.f;
    for (int t = 0; t < 81; ++t) s += fabsf(w1p[t * 32 + c]);
    wsum = fmaxf(wsum, s);
    b1max = fmaxf(b1max, fabsf(b1p[c]));
  }
  const float s1 = pow2_scale_for(m1, 1024.f), s2 = pow2_scale_for(m2, 1024.f);
  W.wsum1 = wsum * 1.0001f;
  W.b1max = b1max;
  W.inv_sw1 = 1.f / s1;
  W.inv_sw2 = 1.f / s2;
  // layer 1: [hi|lo][atom 0..1][32 rows][64]; k = tap (81 used)
  std::vector<__half> img1((size_t)2 * 2 * 32 * 64, __float2half(0.f));
  for (int c = 0; c < 32; ++c)
    for (int t = 0; t < 81; ++t) {
      const float v = w1p[t * 32 + c] * s1;
      const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
      const int atom = t >> 6, k = t & 63;
      img1[(size_t)(0 * 2 + atom) * 32 * 64 + sw128_index(c, k)] = h;
      img1[(size_t)(1 * 2 + atom) * 32 * 64 + sw128_index(c, k)] = l;
    }
  // layer 2: [B tap 0..8][16 rows][64]; row = partial map (ta,tb) (9 used), k = net * 32 + (hi: 0 | lo: 16) + channel
  std::vector<__half> img2((size_t)9 * 16 * 64, __float2half(0.f));
  for (int net = 0; net < 2; ++net)
    for (int d = 0; d < 9; ++d)
      for (int t = 0; t < 9; ++t)
        for (int ch = 0; ch < 16; ++ch) {
          const float v = w2p[(d * 9 + t) * 32 + net * 16 + ch] * s2;
          const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
          img2[(size_t)t * 16 * 64 + sw128_index(d, net * 32 + ch)] = h;
          img2[(size_t)t * 16 * 64 + sw128_index(d, net * 32 + 16 + ch)] = l;
        }
  const size_t b1 = img1.size() * 2, b2 = img2.size() * 2;
  if (W.blob == nullptr) {
    if (cudaMalloc(&W.blob, b1 + b2) != cudaSuccess) {
      cudaGetLastError();
      set_last_error("out of device memory packing the NC weights");
      return -3;
    }
  }
  W.img1 = reinterpret_cast<__half*>(W.blob);
  W.img2 = reinterpret_cast<__half*>(W.blob + b1);
  P2P_CUDA_OK(cudaMemcpy(W.img1, img1.data(), b1, cudaMemcpyHostToDevice));
  P2P_CUDA_OK(cudaMemcpy(W.img2, img2.data(), b2, cudaMemcpyHostToDevice));
  return 0;
}

size_t nc_umma_scratch_bytes(size_t V) { return V * 128 + 18 * V * 4 + 4096; }

// x [hA*wA][hB*wB] -> out (NeighConsensus output); rowmax / colmax (optional) receive the maxima MutualMatching needs.
// xmax: device word holding the float bits of max |x| (launch_absmax or the fused mutual_apply pass).
int launch_neigh_consensus_umma(const float* x, int hA, int wA, int hB, int wB, const NcUmmaWeights& W, const float* b1p,
                                float b2, const unsigned int* xmax, __half* hidden, float* partial, float* out,
                                float* rowmax, unsigned int* colmax, int num_sms, cudaStream_t st) {
  NcParams p;
  memset(&p, 0, sizeof(p));
  p.hA = hA; p.wA = wA; p.hB = hB; p.wB = wB;
  p.nA = hA * wA; p.nB = hB * wB;
  p.V = (long long)p.nA * p.nB;
  p.x = x; p.xmax = xmax; p.hidden = hidden; p.partial = partial; p.b1p = b1p;
  p.wsum1 = W.wsum1; p.b1max = W.b1max; p.inv_sw1 = W.inv_sw1; p.inv_sw2 = W.inv_sw2;
  const long long vt = (p.V + 127) / 128;
  const long long t1 = (long long)p.nA * ((p.nB + 127) / 128);
  P2P_REQUIRE(vt < (1ll << 31) && t1 < (1ll << 31), "NeighConsensus: 4D volume too large");
  {
    p.wimg = W.img1;
    p.tiles = (int)t1;
    const int seg = 9 * nc_l1_rows(wB) * (wB + 2) * 4;
    const int fixed = kL1Stages * 2 * kNcAtom + 2 * 2 * 32 * 128 + 1024;
    p.l1_bufs = fixed + 2 * seg <= 220 * 1024 ? 2 : 1;
    const int smem = fixed + p.l1_bufs * seg;
    P2P_REQUIRE(smem <= 220 * 1024, "NeighConsensus layer 1: B grid too wide for the shared-memory staging (wB <= ~500)");
    auto k = nc_l1_umma_kernel;
    P2P_ENSURE_SMEM(k, smem);
    const int grid = p.tiles < num_sms ? p.tiles : num_sms;
    k<<<grid, kL1Threads, smem, st>>>(p);
    P2P_LAUNCH_OK();
  }
  {
    p.wimg = W.img2;
    p.tiles = (int)vt;
    const int smem = kL2Stages * kNcAtom + 9 * 16 * 128 + 1024;
    auto k = nc_l2_umma_kernel;
    P2P_ENSURE_SMEM(k, smem);
    k<<<p.tiles < num_sms ? p.tiles : num_sms, 512, smem, st>>>(p);
    P2P_LAUNCH_OK();
  }
  if (colmax != nullptr) P2P_CUDA_OK(cudaMemsetAsync(colmax, 0, sizeof(unsigned int) * p.nB, st));
  nc_combine_kernel<<<cdiv(p.nA, kCombineRows), 256, 0, st>>>(partial, hA, wA, p.nB, b2, out, rowmax, colmax);
  P2P_LAUNCH_OK();
  return 0;
}

}  // namespace p2p
