// tcgen05 implementation of the fused MLP paths.  ONE persistent kernel: for every tile of 128 points
// (64 per CTA of a 2-CTA cluster) it runs  [ray -> x_obs -> skinning-voxel sample -> inverse LBS ->]
// positional encoding -> 9 linears (+softplus) -> sdf / 256 features  with every activation kept
// on-chip.  Replaces model/Embedder.py:43-50 + model/network.py:89-119 (and, in render mode, the
// Deformer.py:406-445 gather/blend and a FastMinv inverse) for the forward path.
//
// Mapping to the hardware (B200, sm_100a):
//  * tensor cores: tcgen05.mma.cta_group::2.kind::f16, M = 128 over the CTA pair (64 rows per CTA --
//    full tensor rate, and a 64x512 fp32 layer output takes only 256 of the 512 TMEM columns, so TWO
//    layer accumulators fit and layer l+1's MMAs run while layer l's accumulator is still being drained),
//    N = 256 per instruction (each CTA stages half of every weight tile => weight traffic from L2 is
//    halved), K = 16.  fp32 parity mode (passes = 3): operands are split a = hi + lo in fp16 and every
//    product is 3 MMAs (hi*hi + lo*hi + hi*lo), fp32 accumulation in TMEM.
//  * weights: fp16 hi/lo panels streamed by TMA (128-byte swizzle) from L2 through a 5-slot x 16 KB ring,
//    signalled with mbarrier complete_tx on the leader CTA, slots released by tcgen05.commit.
//  * activations: 64 x 512 fp16 hi + lo (128 KB) in shared memory in the UMMA K-major SW128 layout,
//    rewritten IN PLACE by the epilogue (the full layer output sits in TMEM first), K-block by K-block,
//    each K-block released to the MMA warp through its own mbarrier so the next layer starts as soon as
//    its first 64 inputs exist.
//  * warp roles (22 warps): 0 TMA producer; 1 and 3 MMA issuers of the leader CTA, one per N tile (independent
//    accumulators; barrier polls issued one step ahead of use -- see the note at the issuer code); 2 TMEM
//    allocator; 4-19 epilogue (four warps per TMEM lane quadrant, 16 columns per chunk: tcgen05.ld -> base-2
//    softplus straight from the raw accumulator with the accumulation-bias compensation -> fp16 hi/lo split ->
//    swizzled st.shared); 20-21 prologue for the NEXT tile (point fetch / inverse LBS / PE, or the input rows of
//    the translator / colour networks).
//  * the same engine runs three networks (Net<0> SDF, Net<1> translator + LBS, Net<2> colour) and a forward-mode
//    (value + tangents, 4 rows per point) variant of the first two.
// Every mbarrier wait is bounded: a protocol bug surfaces as a status code, not a hung GPU.
#include "sdf_mlp.cuh"
#include "tc_common.cuh"
#include "../../include/recmv_b200_diag.h"   // recmv_sdf_mlp_tc_debug (diagnostics entry, not in the product header)

namespace recmv {
using namespace tc;

namespace {

#ifndef RECMV_EPI_WARPS
#define RECMV_EPI_WARPS 16
#endif
constexpr int kEpiWarps = RECMV_EPI_WARPS;         // 8 or 16: 2 or 4 warps per TMEM lane quadrant
constexpr int kEpiGroups = kEpiWarps / 4;          // warps sharing a quadrant split every 64-column K block ...
constexpr int kCw = 64 / kEpiGroups;               // ... into chunks of kCw columns
constexpr int kThreads = 32 * (4 + kEpiWarps + 2); // 4 role warps + epilogue warps + 2 prologue warps
static_assert(kEpiWarps == 8 || kEpiWarps == 16, "epilogue warps");
constexpr int kPairs = 1;  // MMA pairs per cluster sharing every weight tile through TMA multicast.
                           // 2 was measured twice (profiles/r01_notes.md): correct and it halves the L2 reads (SM clock
                           // under the power cap rises 1.73 -> 1.85 GHz), but the pairs then consume the weight
                           // stream in lockstep and the frame gets SLOWER (tc3 1.41 -> 1.32 M rays/s, tc1 2.36 -> 2.08);
                           // clusters of 4 also only fit on 132 of the 148 SMs.
constexpr int kRowsPerCta = 64;
#ifndef RECMV_TC_TWO_SWEEPS
#define RECMV_TC_TWO_SWEEPS 1   // last layer, parity mode: correction MMAs first, hi*hi MMAs in a second sweep
#endif
constexpr bool kTwoSweeps = RECMV_TC_TWO_SWEEPS != 0;
constexpr int kSlots = 5;       // weight ring slots in their own region ...
constexpr int kSlotsMax = 9;    // ... + 4 more in the (unused) activation lo plane when a single pass is issued
constexpr uint32_t kSlotBytes = 16384;
// shared memory map (offsets from a 1024-aligned base)
constexpr uint32_t kOffAHi = 0;
constexpr uint32_t kOffALo = 65536;
constexpr uint32_t kOffPeHi = 131072;
constexpr uint32_t kOffPeLo = 139264;
constexpr uint32_t kOffW = 147456;
constexpr uint32_t kOffBar = kOffW + kSlots * kSlotBytes;  // 229376
// barrier indices (8 bytes each)
constexpr int kBarFull = 0;      // [9] pair leader: weight slot filled (1 arming arrival + tx bytes of both CTAs)
constexpr int kBarEmpty = 9;     // [9] cluster CTAs 0/1: slot consumed by every pair (one tcgen05.commit per pair)
constexpr int kBarAReady = 18;   // [8] leader: activation K-block written by both CTAs (4 * kEpiGroups warp arrivals)
constexpr int kBarPeReady = 26;  //     leader: positional-encoding block written (4 warp arrivals)
constexpr int kBarPeFree = 27;   //     local : layer-4 MMAs done with the PE block (commit multicast, both issuers)
constexpr int kBarAccFull = 28;  // [2] local : layer accumulator complete (commit multicast, both issuers)
constexpr int kBarAccEmpty = 30; // [2] leader: accumulator drained by all epilogue warps of the pair
constexpr int kNumBars = 32;
constexpr uint32_t kOffMisc = kOffBar + kNumBars * 8;  // tmem ptr, abort flag, valid flags
constexpr uint32_t kSmemBytes = kOffMisc + 16 + 128 + 1024 /*alignment slack*/;

struct TcParams {
  PointSource src;
  PeWeights pw;
  const float* bias;  // [9][512] padded
  float* out_sdf;
  float* out_feat;
  long long P;
  const int* P_dev;   // optional: actual point count in device memory (<= P)
  int passes;
  float acc_gain_kb;  // relative accumulator gain per 64-wide K block of a layer (4 accumulating MMAs)
  DevStatus* status;
  // diagnostics: raw accumulator (before bias) of layer dbg_layer for tile 0 -> dbg_out [128][512]
  int dbg_layer;
  float* dbg_out;
  // diagnostics: clock64 stamps of cluster 0 / leader CTA, tiles 0-1: trace[role][it][layer][16]
  // role 0 = MMA thread, 1 = first epilogue warp, 2 = last epilogue warp, 3 = producer
  unsigned long long* trace;
  // forward-mode (JVP) variant: out_grad [P,3] = d sdf / d x; each point occupies 4 adjacent tile rows
  float* out_grad;
  // deformer network (kNet == 1): MLPTranslator 167 -> 512 x4 -> 3 (ReLU) followed by LBS forward
  const float* conds;            // [F,128]
  const long long* batch_inds;   // [P] or NULL (frame = p / points_per_frame)
  long long points_per_frame;
  int num_frames;
  const float* bones;            // A [F,24,4,4] (NULL: translator only)
  const float* trans;            // [F,3]
  Voxel vox;
  float* out_translated;         // [P,3]  p + delta
  float* out_offset;             // [P,3]  delta (MLPTranslator.offset), optional
  float* out_posed;              // [P,3]  LBS(p + delta), optional
  float* out_jac;                // [P,9]  forward-mode variant: d out / d p (row i = gradient of output i)
  // colour network (kNet == 2): cat[p, PE4(view), n, feat] = 289 -> 512 x4 (ReLU) -> 3 -> tanh; src.x = points
  const float* normals;          // [P,3]
  const float* view_dirs;        // [P,3]
  const float* feats;            // [P,256]
  float* out_rgb;                // [P,3]
};
#define TRACE(role, it, l, ev)                                                                     \
  do {                                                                                             \
    if (prm.trace && blockIdx.x == 0 && (it) < 2)                                                  \
      prm.trace[((((role) * 2 + (it)) * 9 + (l)) << 4) + (ev)] = (unsigned long long)clock64();     \
  } while (0)

// K-block processing order of a layer's input.  The two warps of a TMEM quadrant split every 64-column
// K block in two 32-column halves, so the epilogue releases K blocks {0,2} (lane halves 0/1 of the first
// N tile), then {1,3}, then {4,6}, {5,7}; the skip layer starts with the PE block, ready long before.
__device__ __forceinline__ int kb_order(int l, int i) {
  if (l == 0) return 0;
  if (l == 4) {
    if (i == 0) return 8;
    --i;
  }
  return (i & 4) | ((i & 1) << 1) | ((i >> 1) & 1);  // 0,2,1,3,4,6,5,7
}

// softplus(beta=100, threshold=20) * kActScale straight from the raw accumulator, in base 2:
//   u = acc * (2^-16 * 100 log2 e) + b * 100 log2 e  (the bias plane is pre-multiplied at pack time)
//   u <= 20 log2 e : ln2/100 * log2(1 + 2^u)     else : z = u / (100 log2 e)        (torch's threshold branch)
// 9 instructions per activation (FFMA FMNMX EX2 FADD LG2 FMUL FMUL FSETP FSEL); ex2/lg2 .ftz skip the denormal
// range fix-ups of __expf/__logf (3 more FMUL + 2 predicates each).
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float softplus100_scaled(uint32_t acc_bits, float bias_u, float kU1 /* acc unscale * 100 log2 e */) {
  constexpr float kUThr = 20.f * 1.4426950408889634f;
  constexpr float kC1 = 0.0069314718055994531f * kActScale;    // ln2 / 100
  constexpr float kC2 = kActScale / kSoftplusLog2Scale;
  const float u = fmaf(__uint_as_float(acc_bits), kU1, bias_u);
  const float y = lg2_ftz(1.f + ex2_ftz(fminf(u, kUThr))) * kC1;
  return u > kUThr ? u * kC2 : y;
}

// value row: softplus100(z); tangent row: z_tangent * sigmoid(100 z_value)  (d softplus / dz; torch's
// threshold branch z > 0.2 has derivative 1, which sigmoid(20+) equals to 2e-9)
__device__ __forceinline__ float act_jvp(float z_own, float z_val, bool is_value) {
  const float t = 100.f * z_val;
  const float e = __expf(-fabsf(t));
  const float inv = __fdividef(1.f, 1.f + e);
  const float sp = t > 20.f ? z_val : (fmaxf(t, 0.f) + __logf(1.f + e)) * 0.01f;
  const float sig = t >= 0.f ? inv : e * inv;
  return is_value ? sp : z_own * sig;
}

// tangent row j of the forward-mode variants: d PE / d x_j  (x -> e_j ; w sin(f x_c) -> w f cos(f x_c) [c == j] ;
// w cos(f x_c) -> -w f sin(f x_c) [c == j])
__device__ __forceinline__ void positional_encode_tangent(float cx, float cy, float cz, int j, const float* pw, float* pe) {
  const float xj = j == 0 ? cx : (j == 1 ? cy : cz);
#pragma unroll
  for (int e = 0; e < 39; ++e) pe[e] = 0.f;
  pe[j] = 1.f;
  float f = 1.f;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float sn, cs;
    sincosf(xj * f, &sn, &cs);
    pe[3 + 6 * k + j] = pw[2 * k] * f * cs;
    pe[3 + 6 * k + 3 + j] = -pw[2 * k + 1] * f * sn;
    f *= 2.f;
  }
}

// Network descriptions driving the same pipeline.
template <int kNet> struct Net;
// SDF network (model/network.py:135-141): PE block -> 512 x3 -> 473 -> [skip: +PE block] 512 x4 -> 257
template <> struct Net<0> {
  static constexpr int kLayers = 9, kPanels = kNumPanels, kInFreeLayer = 4;
  __device__ static int nkb(int l) { return num_panels(l); }
  __device__ static int pbase(int l) { return panel_base(l); }
  __device__ static int kb_at(int l, int i) { return kb_order(l, i); }
  __device__ static bool is_pe(int l, int kbi) { return l == 0 || (l == 4 && kbi == 8); }
  __device__ static int ntiles(int) { return 2; }
  __device__ static bool small(int l, int nt) { return l == 8 && nt == 1; }
};
// MLPTranslator (model/Deformer.py:141-206): [PE(39) | cond(128) | 0 x25] = 3 K blocks -> 512 x4 (ReLU) -> 3
template <> struct Net<1> {
  static constexpr int kLayers = 5, kPanels = 35, kInFreeLayer = 4;
  __device__ static int nkb(int l) { return l == 0 ? 3 : 8; }
  __device__ static int pbase(int l) { return l == 0 ? 0 : 3 + 8 * (l - 1); }
  __device__ static int kb_at(int l, int i) { return l == 0 ? i : ((i & 4) | ((i & 1) << 1) | ((i >> 1) & 1)); }
  __device__ static bool is_pe(int, int) { return false; }   // the input block lives in the activation buffer
  __device__ static int ntiles(int l) { return l == 4 ? 1 : 2; }
  __device__ static bool small(int l, int) { return l == 4; }
};

// RenderingNetwork_view_norm, mode 'idr' (model/RenderNet.py:59-96): [p 3 | PE4(v) 27 | n 3 | feat 256 | 0 x31] = 5 K
// blocks -> 512 x4 (ReLU) -> 3 -> tanh
template <> struct Net<2> {
  static constexpr int kLayers = 5, kPanels = 37, kInFreeLayer = 4;
  __device__ static int nkb(int l) { return l == 0 ? 5 : 8; }
  __device__ static int pbase(int l) { return l == 0 ? 0 : 5 + 8 * (l - 1); }
  __device__ static int kb_at(int l, int i) { return l == 0 ? i : ((i & 4) | ((i & 1) << 1) | ((i >> 1) & 1)); }
  __device__ static bool is_pe(int, int) { return false; }
  __device__ static int ntiles(int l) { return l == 4 ? 1 : 2; }
  __device__ static bool small(int l, int) { return l == 4; }
};

}  // namespace

template <bool kJvp, int kNet>
__global__ void __cluster_dims__(2 * kPairs, 1, 1) __launch_bounds__(kThreads, 1)
sdf_tc_kernel(const __grid_constant__ CUtensorMap tmap128, const TcParams prm) {
  using NetT = Net<kNet>;
  constexpr int kNumLayers = NetT::kLayers;   // shadows the SDF constant of common.cuh
  static_assert(!kJvp || kNet == 0 || kNet == 1, "the forward-mode variant exists for the SDF network and the deformer");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kOffBar;
  volatile int* abort_flag = reinterpret_cast<volatile int*>(gbase + kOffMisc + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffMisc);
  volatile unsigned char* valid = reinterpret_cast<volatile unsigned char*>(gbase + kOffMisc + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();       // rank in the cluster (0 .. 2*kPairs-1)
  const uint32_t pair = crank >> 1;               // MMA pair inside the cluster
  const uint32_t rank = crank & 1u;               // rank inside the pair (0 = leader, issues the MMAs)
  const uint32_t lrank = crank & ~1u;             // cluster rank of this pair's leader
  const bool leader = rank == 0;
  constexpr int kPtsPerTile = kJvp ? 32 : 128;  // JVP: rows = 4 per point (value, d/dx, d/dy, d/dz)
  // number of points: by value, or -- device worklists (coarse-to-fine sweep) -- read from device memory, bounded by
  // the capacity prm.P the grid was sized for
  const long long Ptot = prm.P_dev ? (long long)min((long long)max(__ldg(prm.P_dev), 0), prm.P) : prm.P;
  const long long num_tiles = (Ptot + kPtsPerTile - 1) / kPtsPerTile;
  const long long cluster_id = blockIdx.x / (2 * kPairs), num_clusters = gridDim.x / (2 * kPairs);
  // every pair of a cluster iterates the same number of times (they consume the multicast weight stream
  // in lockstep); iterations past the last tile run on zero rows and write nothing
  const long long n_iter = (num_tiles + (long long)kPairs * num_clusters - 1) / ((long long)kPairs * num_clusters);
  auto tile_of = [&](long long it) { return (cluster_id + it * num_clusters) * kPairs + (long long)pair; };
  const int passes = prm.passes;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  // Single-pass mode never reads the activation lo plane, so its 64 KB can serve as four more weight slots
  // (RECMV_TC_DEEP_RING).  Measured: 9 slots instead of 5 change nothing (tc1 2.32 vs 2.36 M rays/s) -- the weight
  // stream is bound by L2 -> SM bandwidth (~8-9 TB/s aggregate in both modes), not by ring depth x latency.
#ifdef RECMV_TC_DEEP_RING
  const int nslots = passes == 1 ? kSlotsMax : kSlots;
#else
  const int nslots = kSlots;
#endif
  auto SLOT = [&](int s) { return base + (s < kSlots ? kOffW + (uint32_t)s * kSlotBytes : kOffALo + (uint32_t)(s - kSlots) * kSlotBytes); };

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < kSlotsMax; ++s) { mbar_init(BAR(kBarFull + s), 1); mbar_init(BAR(kBarEmpty + s), kPairs); }
    for (int k = 0; k < 8; ++k) mbar_init(BAR(kBarAReady + k), 4 * kEpiGroups);
    mbar_init(BAR(kBarPeReady), 4);
    mbar_init(BAR(kBarPeFree), 2);                               // one commit per MMA issuer
    for (int b = 0; b < 2; ++b) { mbar_init(BAR(kBarAccFull + b), 2); mbar_init(BAR(kBarAccEmpty + b), 2 * kEpiWarps); }
    // arm generation 0 of every weight slot of this pair (both CTAs' halves: 2 x 16 KB)
    if (leader) for (int s = 0; s < kSlotsMax; ++s) mbar_expect_tx_local(BAR(kBarFull + s), 2 * kSlotBytes);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) prefetch_tmap(&tmap128);
  if (warp == 2) tmem_alloc_pair(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============== TMA producer: cluster CTAs 0 and 1 load one half each, multicast to all pairs ==============
    if (lane == 0 && crank < 2) {
      int slot = 0;
      uint32_t ring = 0;
      uint16_t mask = 0;
      for (int g = 0; g < kPairs; ++g) mask |= (uint16_t)(1u << (2 * g + (int)rank));
      for (long long it = 0; it < n_iter; ++it) {
        for (int l = 0; l < kNumLayers; ++l) {
          const int nkb = NetT::nkb(l);
          TRACE(3, it, l, 0);
          // last layer, parity mode: sweep 0 streams (w_hi, w_lo) for the correction MMAs, sweep 1 streams
          // w_hi again for the hi*hi MMAs (see the MMA issuer)
          const int nsweep = (kTwoSweeps && passes == 3 && l == kNumLayers - 1) ? 2 : 1;
          for (int sweep = 0; sweep < nsweep; ++sweep) {
            for (int i = 0; i < nkb; ++i) {
              const int kbi = NetT::kb_at(l, i);
              if (i == nkb - 1) TRACE(3, it, l, 1);
              for (int nt = 0; nt < NetT::ntiles(l); ++nt) {
                const int nplanes = (passes == 3 && sweep == 0) ? 2 : 1;
                for (int plane = 0; plane < nplanes; ++plane) {
                  mbar_wait(BAR(kBarEmpty + slot), ring ^ 1u, abort_flag, prm.status, 100 + slot);
                  const int row = (plane * NetT::kPanels + NetT::pbase(l) + kbi) * 512 + nt * 256 + (int)rank * 128;
                  tma_load_2d_pair_mcast(SLOT(slot), (const void*)&tmap128,
                                         mapa(BAR(kBarFull + slot), 0), mask, 0, row);
                  if (++slot == nslots) { slot = 0; ring ^= 1u; }
                }
              }
            }
          }
        }
      }
      // drain: the releases of the last `nslots` slot uses are asynchronous tcgen05.commit arrivals that no load
      // waits for any more; collect them before this CTA may exit (same hazard as the final pe_free, see below)
      for (int k = 0; k < nslots; ++k) {
        mbar_wait(BAR(kBarEmpty + slot), ring ^ 1u, abort_flag, prm.status, 110 + slot);
        if (++slot == nslots) { slot = 0; ring ^= 1u; }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ====================== MMA issuers (leader CTA only): warp 1 -> N tile 0, warp 3 -> N tile 1 ===============
    // Measured with tools/tc_microbench.py (profiles/r01_tc_microbench.txt): one thread that also polls the
    // full / ready barriers and commits per slot feeds the 64-cycle M128 x N256 MMAs at only ~125-165
    // cycles each once the epilogue warps compete for its scheduler; two issuing threads (independent
    // accumulators, so no ordering between them is needed) with their barrier polls issued one step AHEAD
    // of use keep the tensor pipe at 64 cycles / MMA.
    if (leader && lane == 0) {
      const int my_nt = warp == 1 ? 0 : 1;
      int slot = 0;          // ring position of the next slot use in producer order
      uint32_t ring = 0;
      auto advance = [&](int n) { slot += n; if (slot >= nslots) { slot -= nslots; ring ^= 1u; } };  // n < nslots
      // try_wait results obtained ahead of time: {barrier, parity, completed}
      uint32_t fh_bar = 0, fh_par = 0, fh_ok = 0, ah_bar = 0, ah_par = 0, ah_ok = 0;
      auto poll = [&](uint32_t bar, uint32_t par) -> uint32_t { return mbar_try_wait(bar, par) ? 1u : 0u; };
#ifdef RECMV_TC_WAITSTATS
      long long ws_full = 0, ws_nblock = 0, ws_a = 0;
#endif
      auto wait_full = [&](int tag) {
        const uint32_t bar = BAR(kBarFull + slot);
#ifdef RECMV_TC_WAITSTATS
        const long long w0 = clock64();
        if (!(fh_ok && fh_bar == bar && fh_par == ring)) { mbar_wait(bar, ring, abort_flag, prm.status, tag + slot); ++ws_nblock; }
        ws_full += clock64() - w0;
#else
        if (!(fh_ok && fh_bar == bar && fh_par == ring)) mbar_wait(bar, ring, abort_flag, prm.status, tag + slot);
#endif
        fh_ok = 0;
        mbar_expect_tx_local(bar, 2 * kSlotBytes);  // arm the slot's next generation
        tc_fence_after();
      };
      auto poll_full_at = [&](int ahead) {   // poll the slot `ahead` uses after the current position
        int s2 = slot + ahead; uint32_t r2 = ring;
        if (s2 >= nslots) { s2 -= nslots; r2 ^= 1u; }
        fh_bar = BAR(kBarFull + s2); fh_par = r2; fh_ok = poll(fh_bar, r2);
      };
      const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
      for (long long it = 0; it < n_iter; ++it) {
        for (int l = 0; l < kNumLayers; ++l) {
          const long long L = it * kNumLayers + l;
          const int buf = (int)(L & 1);
          const uint32_t use = (uint32_t)(L >> 1);
          if (my_nt == 0) TRACE(0, it, l, 0);
          mbar_wait(BAR(kBarAccEmpty + buf), (use & 1u) ^ 1u, abort_flag, prm.status, 200 + buf);
          tc_fence_after();
          if (my_nt == 0) TRACE(0, it, l, 1);
          const int nkb = NetT::nkb(l);
          const int T = NetT::ntiles(l);
          const bool small = NetT::small(l, my_nt);
          const uint32_t idesc = small ? idesc_f16(128, 32) : idesc_f16(128, 256);
          const uint32_t dcol = tmem_base + (uint32_t)(buf * 256 + my_nt * 128);
          // Ordering for accuracy (parity mode, last layer): the tensor core accumulates with truncation, so
          // every MMA added to a LARGE accumulator costs ~1 ulp of it.  The correction products (lo*hi,
          // hi*lo) are 2^-11 of the result: issued first, while the accumulator is still tiny, their
          // truncations are negligible; the 32 hi*hi MMAs follow in a second sweep over the K blocks
          // (96 -> 32 significant truncations).  Other layers keep one sweep (w_hi is streamed once).
          const bool two_sweeps = (kTwoSweeps && passes == 3 && l == kNumLayers - 1);
          for (int sweep = 0; sweep < (two_sweeps ? 2 : 1); ++sweep) {
            const int planes = (passes == 3 && sweep == 0) ? 2 : 1;   // slot uses per (K block, tile)
            for (int i = 0; i < nkb; ++i) {
              if (my_nt >= T) { advance(T * planes); continue; }      // single-tile layer: nothing for issuer 1
              const int kbi = NetT::kb_at(l, i);
              const bool is_pe = NetT::is_pe(l, kbi);
              if (sweep == 0) {
                // layer 0 consumes what the prologue wrote (PE buffer, or the input K blocks of the deformer);
                // every other layer the K blocks released by the previous layer's epilogue
                if (l == 0 || !is_pe) {
                  const uint32_t bar = l == 0 ? BAR(kBarPeReady) : BAR(kBarAReady + kbi);
                  const uint32_t par = l == 0 ? (uint32_t)(it & 1) : (uint32_t)((l - 1) & 1);
#ifdef RECMV_TC_WAITSTATS
                  const long long w0 = clock64();
#endif
                  if (!(ah_ok && ah_bar == bar && ah_par == par)) mbar_wait(bar, par, abort_flag, prm.status, 220 + kbi);
                  ah_ok = 0;
#ifdef RECMV_TC_WAITSTATS
                  ws_a += clock64() - w0;
#endif
                }
                tc_fence_after();
                if (my_nt == 0) {
                  if (i == 0) TRACE(0, it, l, 2);
                  if (i == 4) TRACE(0, it, l, 3);
                  if (i == nkb - 1) TRACE(0, it, l, 4);
                }
              }
              const uint64_t a_hi = smem_desc_sw128(is_pe ? base + kOffPeHi : base + kOffAHi + kbi * 8192);
              const uint64_t a_lo = smem_desc_sw128(is_pe ? base + kOffPeLo : base + kOffALo + kbi * 8192);
              advance(my_nt * planes);                                // the lower tile's uses of this K block
              // ---- slot with w_hi ---------------------------------------------------------------------------
              wait_full(230);
              uint64_t b = smem_desc_sw128(SLOT(slot));
              const int s_hi = slot;
              advance(1);
              // poll ahead: this K block's w_lo slot, or the next K block's first slot
              poll_full_at(planes == 2 ? 0 : (T - 1) * planes);
              if (sweep == 0 && l > 0 && i + 1 < nkb) {               // ... and the next K block's activations
                const int kb2 = NetT::kb_at(l, i + 1);
                if (!NetT::is_pe(l, kb2)) { ah_bar = BAR(kBarAReady + kb2); ah_par = (uint32_t)((l - 1) & 1); ah_ok = poll(ah_bar, ah_par); }
              }
              if (!two_sweeps) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_pair(dcol, a_hi + 2 * k, b + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
                if (passes == 3) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) umma_f16_pair(dcol, a_lo + 2 * k, b + 2 * k, idesc, 1u);
                }
              } else if (sweep == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_pair(dcol, a_lo + 2 * k, b + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_pair(dcol, a_hi + 2 * k, b + 2 * k, idesc, 1u);
              }
              umma_commit_pair(BAR(kBarEmpty + s_hi), 3);  // cluster CTAs 0 and 1 (the producers)
              // ---- slot with w_lo (parity mode; not in the hi*hi sweep) --------------------------------------
              if (planes == 2) {
                wait_full(240);
                b = smem_desc_sw128(SLOT(slot));
                const int s_lo = slot;
                advance(1);
                poll_full_at((T - 1) * planes);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_pair(dcol, a_hi + 2 * k, b + 2 * k, idesc, 1u);
                umma_commit_pair(BAR(kBarEmpty + s_lo), 3);
              }
              advance((T - 1 - my_nt) * planes);                      // the upper tile's uses of this K block
            }
          }
          umma_commit_pair(BAR(kBarAccFull + buf), pair_mask);   // one arrival per issuer
          if (my_nt == 0) TRACE(0, it, l, 5);
#ifdef RECMV_TC_WAITSTATS
          if (prm.trace && blockIdx.x == 0 && it < 2) {   // per layer: cycles in weight-slot waits, blocking waits, K-block waits
            unsigned long long* t = prm.trace + (((0 * 2 + it) * 9 + l) << 4) + 6 + 3 * my_nt;
            t[0] = (unsigned long long)ws_full; t[1] = (unsigned long long)ws_nblock; t[2] = (unsigned long long)ws_a;
          }
          ws_full = ws_nblock = ws_a = 0;
#endif
          if (l == NetT::kInFreeLayer) umma_commit_pair(BAR(kBarPeFree), pair_mask);  // prologue may refill its block
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ================================ epilogue (both CTAs) =========================================
    const int q = warp & 3;               // TMEM lane quadrant == warp index % 4
    const int grp = (warp - 4) >> 2;      // which kCw-column part of every K block of this quadrant
    const int row = (q & 1) * 32 + lane;  // tile row owned by this thread (lanes 64.. mirror rows 0..63)
    const int half = q >> 1;              // which 128-column half of each 256-wide N tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (long long it = 0; it < n_iter; ++it) {
      const long long tile = tile_of(it);
      const long long p = kJvp ? tile * 32 + (long long)rank * 16 + (row >> 2)
                               : tile * 128 + (long long)rank * kRowsPerCta + row;
      const int comp = kJvp ? (row & 3) : 0;   // 0 = value row, 1..3 = tangent rows
      const bool is_value = comp == 0;
      for (int l = 0; l < kNumLayers; ++l) {
        const long long L = it * kNumLayers + l;
        const int buf = (int)(L & 1);
        const uint32_t use = (uint32_t)(L >> 1);
        const float* bias = prm.bias + l * 512;
        const bool hidden = l < kNumLayers - 1;
        // 2^-16 x (1 + compensation of the tensor core's truncating accumulation: every one of the K/16 hi*hi
        // MMAs that adds into the full-size fp32 accumulator drops on average ~2^-24 of it; recmv_tc_set_acc_gain)
        const float acc_unscale = kAccUnscale * (1.f + prm.acc_gain_kb * (float)NetT::nkb(l));
        const float acc_u1 = acc_unscale * kSoftplusLog2Scale;
        // hidden layers: 4 chunks of kCw columns per warp, chunk c = (N tile c >> 1, K block c & 1 of this half)
        const float* bsel = (!kJvp && kNet == 0) ? bias + kNumLayers * 512 : bias;   // SDF: the b * 100 log2 e plane
        auto f0_of = [&](int c) { return (c >> 1) * 256 + half * 128 + (c & 1) * 64 + grp * kCw; };
        float bcur[kCw];
        if (hidden) {   // the first chunk's biases travel while this warp waits for the accumulator
#pragma unroll
          for (int j = 0; j < kCw / 4; ++j) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(bsel + f0_of(0) + 4 * j));
            bcur[4 * j] = t.x; bcur[4 * j + 1] = t.y; bcur[4 * j + 2] = t.z; bcur[4 * j + 3] = t.w;
          }
        }
        const bool tr = lane == 0 && (warp == 4 || warp == 3 + kEpiWarps);
        const int trole = warp == 4 ? 1 : 2;
        if (tr) TRACE(trole, it, l, 0);
        mbar_wait(BAR(kBarAccFull + buf), use & 1u, abort_flag, prm.status, 300 + q);
        tc_fence_after();
        if (tr) TRACE(trole, it, l, 1);
        if (hidden) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int f0 = f0_of(c);
            uint32_t r[kCw];
            tmem_ld_cols<kCw>(tmem_base + lane_addr + (uint32_t)(buf * 256 + (c >> 1) * 128 + (c & 1) * 64 + grp * kCw), r);
            float bnext[kCw];
            if (c < 3) {   // next chunk's biases: in flight during this chunk's arithmetic
#pragma unroll
              for (int j = 0; j < kCw / 4; ++j) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(bsel + f0_of(c + 1) + 4 * j));
                bnext[4 * j] = t.x; bnext[4 * j + 1] = t.y; bnext[4 * j + 2] = t.z; bnext[4 * j + 3] = t.w;
              }
            }
            tmem_ld_wait();
            if (c == 0 && tr) TRACE(trole, it, l, 5);
            if (prm.dbg_out && prm.dbg_layer == l && tile == 0) {  // cluster 0, pair 0
              float* d = prm.dbg_out + ((size_t)rank * kRowsPerCta + row) * 512 + f0;
#pragma unroll
              for (int cc = 0; cc < kCw; ++cc) d[cc] = __uint_as_float(r[cc]) * kAccUnscale;
            }
#pragma unroll
            for (int j = 0; j < kCw / 8; ++j) {
              const int f = f0 + 8 * j;
              float v[8];
              if (!kJvp && kNet == 0) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = softplus100_scaled(r[8 * j + e], bcur[8 * j + e], acc_u1);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaf(__uint_as_float(r[8 * j + e]), acc_unscale, bcur[8 * j + e]);
                if (kJvp) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float z_own = is_value ? v[e] : __uint_as_float(r[8 * j + e]) * acc_unscale;  // tangents: no bias
                    const float z_val = __shfl_sync(0xffffffffu, z_own, lane & ~3);  // the point's value row
                    v[e] = (kNet == 0 ? act_jvp(z_own, z_val, is_value) : (z_val > 0.f ? z_own : 0.f)) * kActScale;
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f) * kActScale;   // ReLU (deformer)
                }
              }
              range_check8(v, prm.status, 1000 + l);
              uint4 hi, lo;
              split8(v, hi, lo);
              const int kb = f >> 6, chunk = (f & 63) >> 3;
              const uint32_t off = (uint32_t)kb * 8192u + sw128_offset(row, chunk);
              st_shared_v4(base + kOffAHi + off, hi);
              if (passes == 3) st_shared_v4(base + kOffALo + off, lo);
            }
            if (c == 0 && tr) TRACE(trole, it, l, 6);
            // this warp's kCw columns of K block (f0 >> 6) of the next layer's input are complete
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(BAR(kBarAReady + (f0 >> 6)), lrank);
            if (tr) TRACE(trole, it, l, c == 0 ? 7 : 2 + (c >> 1));
            if (c < 3) {
#pragma unroll
              for (int e = 0; e < kCw; ++e) bcur[e] = bnext[e];
            }
          }
        }
        for (int nt = 0; !hidden && nt < NetT::ntiles(l); ++nt) {
          const bool small = NetT::small(l, nt);
          if (small && grp != 0) continue;  // the 32-wide tail tile has a single 16-column group per half
          for (int kbl = 0; kbl < (small ? 1 : 2); ++kbl) {
            const int c0 = small ? 0 : kbl * 64 + grp * kCw;  // K block kbl of this 128-column half, part grp
            uint32_t r[kCw];
            tmem_ld_cols<kCw>(tmem_base + lane_addr + (uint32_t)(buf * 256 + nt * 128 + c0), r);
            tmem_ld_wait();
            const int f0 = small ? nt * 256 + half * 16 + c0 : nt * 256 + half * 128 + c0;
            if (prm.dbg_out && prm.dbg_layer == l && tile == 0 && !small) {  // cluster 0, pair 0
              float* d = prm.dbg_out + ((size_t)rank * kRowsPerCta + row) * 512 + f0;
#pragma unroll
              for (int c = 0; c < kCw; ++c) d[c] = __uint_as_float(r[c]) * kAccUnscale;
            }
            if (kNet == 2) {
              // colour network: columns 0..2 of the tail tile -> tanh
              if (p < Ptot && half == 0 && c0 == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                  prm.out_rgb[3 * p + c] = tanhf(fmaf(__uint_as_float(r[c]), acc_unscale, __ldg(bias + c)));
              }
            } else if (kNet == 1) {
              // deformer: columns 0..2 of the tail tile = offset; out = p + offset, then LBS forward
              if (kJvp) {
                if (half == 0 && c0 == 0) {   // warp-uniform; rows 4i..4i+3 = (value, d/dx, d/dy, d/dz) in adjacent lanes
                  float dv[3], dt[9];         // offset and d offset_c / d p_j
#pragma unroll
                  for (int c = 0; c < 3; ++c) {
                    const float own = __uint_as_float(r[c]) * acc_unscale;
                    dv[c] = own + __ldg(bias + c);
#pragma unroll
                    for (int j = 0; j < 3; ++j) dt[3 * c + j] = __shfl_sync(0xffffffffu, own, (lane & ~3) + 1 + j);
                  }
                  if (is_value && p < Ptot) {
                    const float x = __ldg(prm.src.x + 3 * p), y = __ldg(prm.src.x + 3 * p + 1), z = __ldg(prm.src.x + 3 * p + 2);
                    const float tx = x + dv[0], ty = y + dv[1], tz = z + dv[2];
                    dt[0] += 1.f; dt[4] += 1.f; dt[8] += 1.f;   // d (p + offset) / d p
                    if (prm.out_translated) { prm.out_translated[3 * p] = tx; prm.out_translated[3 * p + 1] = ty; prm.out_translated[3 * p + 2] = tz; }
                    if (prm.out_offset) { prm.out_offset[3 * p] = dv[0]; prm.out_offset[3 * p + 1] = dv[1]; prm.out_offset[3 * p + 2] = dv[2]; }
                    if (prm.out_posed) {
                      long long f = prm.batch_inds ? prm.batch_inds[p] : (prm.points_per_frame > 0 ? p / prm.points_per_frame : 0);
                      f = f < 0 ? 0 : (f >= prm.num_frames ? prm.num_frames - 1 : f);
                      float o[3], Jl[9];
                      lbs_forward_jac(prm.vox, prm.bones + (size_t)f * 384, prm.trans + 3 * f, tx, ty, tz, o, Jl);
#pragma unroll
                      for (int i = 0; i < 3; ++i) {
                        prm.out_posed[3 * p + i] = o[i];
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                          prm.out_jac[9 * p + 3 * i + j] = Jl[3 * i] * dt[j] + Jl[3 * i + 1] * dt[3 + j] + Jl[3 * i + 2] * dt[6 + j];
                      }
                    } else {
#pragma unroll
                      for (int e = 0; e < 9; ++e) prm.out_jac[9 * p + e] = dt[e];
                    }
                  }
                }
              } else if (p < Ptot && half == 0 && c0 == 0) {
                const float dx = fmaf(__uint_as_float(r[0]), acc_unscale, __ldg(bias + 0));
                const float dy = fmaf(__uint_as_float(r[1]), acc_unscale, __ldg(bias + 1));
                const float dz = fmaf(__uint_as_float(r[2]), acc_unscale, __ldg(bias + 2));
                const float x = __ldg(prm.src.x + 3 * p), y = __ldg(prm.src.x + 3 * p + 1), z = __ldg(prm.src.x + 3 * p + 2);
                const float tx = x + dx, ty = y + dy, tz = z + dz;
                if (prm.out_translated) { prm.out_translated[3 * p] = tx; prm.out_translated[3 * p + 1] = ty; prm.out_translated[3 * p + 2] = tz; }
                if (prm.out_offset) { prm.out_offset[3 * p] = dx; prm.out_offset[3 * p + 1] = dy; prm.out_offset[3 * p + 2] = dz; }
                if (prm.out_posed) {
                  long long f = prm.batch_inds ? prm.batch_inds[p] : (prm.points_per_frame > 0 ? p / prm.points_per_frame : 0);
                  f = f < 0 ? 0 : (f >= prm.num_frames ? prm.num_frames - 1 : f);
                  float w[24], T[12];
                  sample_skin24(prm.vox, tx, ty, tz, w);
                  blend_bones(prm.bones + (size_t)f * 384, w, T);
                  const float* tr = prm.trans + 3 * f;
                  prm.out_posed[3 * p + 0] = (T[0] * tx + T[1] * ty + T[2] * tz + T[3]) + __ldg(tr + 0);
                  prm.out_posed[3 * p + 1] = (T[4] * tx + T[5] * ty + T[6] * tz + T[7]) + __ldg(tr + 1);
                  prm.out_posed[3 * p + 2] = (T[8] * tx + T[9] * ty + T[10] * tz + T[11]) + __ldg(tr + 2);
                }
              }
            } else if (p < Ptot) {
              // last layer: column 0 = sdf, columns 1..256 = features
              const bool ok = valid[(it & 1) * 64 + row] != 0;
              if (kJvp && !is_value) {
                if (!small && f0 == 0) prm.out_grad[p * 3 + (comp - 1)] = __uint_as_float(r[0]) * acc_unscale;
              } else if (!small) {
                if (f0 == 0) prm.out_sdf[p] = ok ? fmaf(__uint_as_float(r[0]), acc_unscale, __ldg(bias)) : kInvalidSdf;
                if (prm.out_feat) {
#pragma unroll
                  for (int c = 0; c < kCw; ++c) {
                    const int f = f0 + c;
                    if (f > 0) prm.out_feat[p * 256 + (f - 1)] = fmaf(__uint_as_float(r[c]), acc_unscale, __ldg(bias + f));
                  }
                }
              } else if (half == 0 && c0 == 0 && prm.out_feat) {
                prm.out_feat[p * 256 + 255] = fmaf(__uint_as_float(r[0]), acc_unscale, __ldg(bias + 256));
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(BAR(kBarAccEmpty + buf), lrank);
        if (tr) TRACE(trole, it, l, 4);
      }
    }
  } else if (warp >= 4 + kEpiWarps) {
    // ================================ prologue for the next tile (both CTAs) ========================
    const int row = (warp - 4 - kEpiWarps) * 32 + lane;
    for (long long it = 0; it < n_iter; ++it) {
      const long long tile = tile_of(it);
      const long long p = kJvp ? tile * 32 + (long long)rank * 16 + (row >> 2)
                               : tile * 128 + (long long)rank * kRowsPerCta + row;
      bool ok = true;
      if (kNet == 2) {
        // colour-network input row, 8 columns at a time straight into the activation buffer (K blocks 0..4)
        float head[40];   // [p 3 | PE4(v) 27 | n 3 | feat 0..6]
#pragma unroll
        for (int e = 0; e < 40; ++e) head[e] = 0.f;
        const bool live = p < Ptot;
        const float* ft = prm.feats + (size_t)(live ? p : 0) * 256;
        if (live) {
          float pe[39];
          positional_encode(__ldg(prm.view_dirs + 3 * p), __ldg(prm.view_dirs + 3 * p + 1), __ldg(prm.view_dirs + 3 * p + 2),
                            prm.pw.w, pe);   // the first 27 entries of a 6-band encoding are the 4-band encoding
#pragma unroll
          for (int e = 0; e < 3; ++e) { head[e] = __ldg(prm.src.x + 3 * p + e); head[30 + e] = __ldg(prm.normals + 3 * p + e); }
#pragma unroll
          for (int e = 0; e < 27; ++e) head[3 + e] = pe[e];
#pragma unroll
          for (int e = 0; e < 7; ++e) head[33 + e] = __ldg(ft + e);
        }
        if (it > 0) mbar_wait(BAR(kBarPeFree), (uint32_t)((it - 1) & 1), abort_flag, prm.status, 400);
#pragma unroll 1
        for (int chunk = 0; chunk < 40; ++chunk) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = 8 * chunk + e;              // feature j sits in column 33 + j
            v[e] = (live && col >= 40 && col < 289) ? __ldg(ft + (col - 33)) : 0.f;
          }
          if (chunk < 5) {
#pragma unroll
            for (int c5 = 0; c5 < 5; ++c5)
              if (chunk == c5) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = head[8 * c5 + e];
              }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= kActScale;
          range_check8(v, prm.status, 1100);
          uint4 hi, lo;
          split8(v, hi, lo);
          const uint32_t off = (uint32_t)(chunk >> 3) * 8192u + sw128_offset(row, chunk & 7);
          st_shared_v4(base + kOffAHi + off, hi);
          if (passes == 3) st_shared_v4(base + kOffALo + off, lo);
        }
      } else if (kNet == 1) {
        // deformer input row: [PE(p) (39) | cond[frame] (128) | 0 (25)] = three 64-wide K blocks written into the
        // activation buffer itself (free again once the previous tile's last layer has been issued and done)
        // (written 8 columns at a time; only the 40-wide head [PE | first cond value] lives in registers)
        float head[40];
#pragma unroll
        for (int e = 0; e < 40; ++e) head[e] = 0.f;
        const bool live = p < Ptot;
        const bool value_row = !kJvp || (row & 3) == 0;
        const float* cond = prm.conds;
        if (live) {
          const float x = __ldg(prm.src.x + 3 * p), y = __ldg(prm.src.x + 3 * p + 1), z = __ldg(prm.src.x + 3 * p + 2);
          if (value_row) {
            positional_encode(x, y, z, prm.pw.w, head);
            long long f = prm.batch_inds ? prm.batch_inds[p] : (prm.points_per_frame > 0 ? p / prm.points_per_frame : 0);
            f = f < 0 ? 0 : (f >= prm.num_frames ? prm.num_frames - 1 : f);
            cond = prm.conds + (size_t)f * 128;
            head[39] = __ldg(cond);
          } else {
            positional_encode_tangent(x, y, z, (row & 3) - 1, prm.pw.w, head);   // the condition does not depend on p
          }
#pragma unroll
          for (int e = 0; e < 40; ++e) head[e] *= kActScale;
        }
        if (it > 0) mbar_wait(BAR(kBarPeFree), (uint32_t)((it - 1) & 1), abort_flag, prm.status, 400);
#pragma unroll
        for (int chunk = 0; chunk < 5; ++chunk) {          // columns 0..39
          range_check8(head + 8 * chunk, prm.status, 1101);
          uint4 hi, lo;
          split8(head + 8 * chunk, hi, lo);
          const uint32_t off = sw128_offset(row, chunk);
          st_shared_v4(base + kOffAHi + off, hi);
          if (passes == 3) st_shared_v4(base + kOffALo + off, lo);
        }
#pragma unroll 1
        for (int chunk = 5; chunk < 24; ++chunk) {         // columns 40..191: cond[1..127] then zeros
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int ci = 8 * chunk + e - 39;             // index into the condition vector
            v[e] = (live && value_row && ci < 128) ? __ldg(cond + ci) * kActScale : 0.f;
          }
          range_check8(v, prm.status, 1102);
          uint4 hi, lo;
          split8(v, hi, lo);
          const uint32_t off = (uint32_t)(chunk >> 3) * 8192u + sw128_offset(row, chunk & 7);
          st_shared_v4(base + kOffAHi + off, hi);
          if (passes == 3) st_shared_v4(base + kOffALo + off, lo);
        }
      } else {
      float pe[40];
      if (p < Ptot) {
        float cx, cy, cz;
        ok = fetch_point(prm.src, p, cx, cy, cz);
        if (!kJvp || (row & 3) == 0) {
          positional_encode(cx, cy, cz, prm.pw.w, pe);
        } else {
          positional_encode_tangent(cx, cy, cz, (row & 3) - 1, prm.pw.w, pe);
        }
#pragma unroll
        for (int e = 0; e < 39; ++e) pe[e] *= kActScale;
      } else {
#pragma unroll
        for (int e = 0; e < 39; ++e) pe[e] = 0.f;
      }
      pe[39] = 0.f;
      if (it > 0) mbar_wait(BAR(kBarPeFree), (uint32_t)((it - 1) & 1), abort_flag, prm.status, 400);
      const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int chunk = 0; chunk < 8; ++chunk) {
        uint4 hi = zero, lo = zero;
        if (chunk < 5) { range_check8(pe + 8 * chunk, prm.status, 1103); split8(pe + 8 * chunk, hi, lo); }
        const uint32_t off = sw128_offset(row, chunk);
        st_shared_v4(base + kOffPeHi + off, hi);
        st_shared_v4(base + kOffPeLo + off, lo);
      }
      }
      valid[(it & 1) * 64 + row] = ok ? 1 : 0;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(BAR(kBarPeReady), lrank);
    }
    // The pe_free commit of the LAST tile is an asynchronous arrival (it fires when the tensor pipe retires the
    // MMAs) that nobody consumes; for the deformer / colour networks it is issued after the final acc_full commit.
    // Wait for it here so that it cannot land after this CTA has exited -- in the shared memory of the next
    // launch's CTA (observed as a sporadic `unspecified launch failure` in back-to-back launches).
    if (n_iter > 0) mbar_wait(BAR(kBarPeFree), (uint32_t)((n_iter - 1) & 1), abort_flag, prm.status, 401);
  }

  // ---- teardown: every role has drained; both CTAs must be done before TMEM goes away ---------------
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

namespace {
DevStatus* g_status_host[16] = {nullptr};
// relative compensation of the accumulation bias per 64-wide K block (= 4 accumulating hi*hi MMAs x 2^-24), per
// precision mode {3 passes, 1 pass}; calibrated against the reference's fp32 results (tools/calibrate_acc_gain.py)
float g_acc_gain_kb[2] = {4.f * 5.9604645e-8f, 4.f * 5.9604645e-8f};
inline void set_acc_scales(TcParams& prm) { prm.acc_gain_kb = g_acc_gain_kb[prm.passes == 3 ? 0 : 1]; }

// one mapped, pinned status record per device
int tc_status_record(int dev, DevStatus** out);
}  // namespace
int device_status_record(void** out) {
  int dev = 0;
  cudaGetDevice(&dev);
  DevStatus* sd = nullptr;
  int s = tc_status_record(dev, &sd);
  *out = sd;
  return s;
}
namespace {
int tc_status_record(int dev, DevStatus** out) {
  DevStatus*& h = g_status_host[dev & 15];
  if (!h) {
    cudaError_t e = cudaHostAlloc((void**)&h, sizeof(DevStatus), cudaHostAllocMapped);
    if (e != cudaSuccess) return (int)e;
    h->code = 0; h->detail = 0; h->block = 0; h->pad = 0;
  }
  *out = h;
  return 0;
}

struct TmapCacheEntry {
  const void* base = nullptr;
  CUtensorMap m128;
};
using TmapCache = TmapCacheEntry;

// opt-in to the 226 KB of dynamic shared memory once per device for every instantiation
int tc_prepare_launch(int dev) {
  static bool done[16] = {false};
  if (done[dev & 15]) return 0;
  cudaError_t e = cudaFuncSetAttribute(sdf_tc_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(sdf_tc_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(sdf_tc_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(sdf_tc_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(sdf_tc_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  done[dev & 15] = true;
  return 0;
}

// clusters must fit inside a GPC: ask the driver how many are co-resident and run exactly that many
// (persistent kernel; a second wave would double the time)
int tc_max_clusters(int dev) {
  static int cached[16] = {0};
  if (cached[dev & 15]) return cached[dev & 15];
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_sms() / (2 * kPairs) * (2 * kPairs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2 * kPairs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, sdf_tc_kernel<false, 0>, &cfg);
  if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms() / (2 * kPairs) - 4; }
  cached[dev & 15] = n;
  return n;
}
}  // namespace

static int launch_tc(const PointSource& src, const void* packed, const PeWeights& pw, float* out_sdf,
                     float* out_feat, float* out_grad, int64_t P, int passes, int dbg_layer, float* dbg_out, int* status_host,
                     unsigned long long* trace, cudaStream_t st, const int* P_dev = nullptr) {
  if (passes != 1 && passes != 3) return RECMV_E_DTYPE;
  PackedLayout L = packed_layout();
  const char* pb = (const char*)packed;
  static TmapCache cache[16];
  int dev = 0;
  cudaGetDevice(&dev);
  TmapCache& tc_ = cache[dev & 15];
  const void* panels = pb + L.f16_off;
  if (tc_.base != panels) {
    int s = make_panel_tmap(&tc_.m128, panels, (uint64_t)2 * kNumPanels * 512, 128);
    if (s) return s;
    tc_.base = panels;
  }
  int sp = tc_prepare_launch(dev);
  if (sp) return sp;
  // Device status record in MAPPED pinned host memory: a bounded wait that times out writes it there, the
  // host sees it without a synchronisation and refuses further work (recmv_check_async_errors()).
  DevStatus* sd = nullptr;
  int s0 = tc_status_record(dev, &sd);
  if (s0) return s0;
  if (sd->code != 0) return RECMV_E_DEVICE;  // a previous launch on this device aborted
  TcParams prm = {};
  prm.src = src; prm.pw = pw;
  prm.bias = (const float*)(pb + L.bias_all_off);
  prm.out_sdf = out_sdf; prm.out_feat = out_feat; prm.P = P; prm.P_dev = P_dev; prm.passes = passes; prm.status = sd;
  set_acc_scales(prm);
  prm.dbg_layer = dbg_layer; prm.dbg_out = dbg_out; prm.trace = trace; prm.out_grad = out_grad;
  const int pts_per_tile = out_grad ? 32 : 128;
  int64_t tiles = (P + pts_per_tile - 1) / pts_per_tile;
  int64_t want = (tiles + kPairs - 1) / kPairs;
  const int maxc = tc_max_clusters(dev);
  int clusters = (int)(want < maxc ? want : maxc);
  if (out_grad)
    sdf_tc_kernel<true, 0><<<clusters * 2 * kPairs, kThreads, kSmemBytes, st>>>(tc_.m128, prm);
  else
    sdf_tc_kernel<false, 0><<<clusters * 2 * kPairs, kThreads, kSmemBytes, st>>>(tc_.m128, prm);
  int s = launch_status();
  if (s) return s;
  if (status_host) {  // diagnostics path: wait for the launch, then report (and clear) the status record
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    status_host[0] = sd->code; status_host[1] = sd->detail; status_host[2] = sd->block; status_host[3] = 0;
    sd->code = 0; sd->detail = 0; sd->block = 0;
  }
  return RECMV_OK;
}

int tc_sdf_forward(const PointSource& src, const void* packed, const PeWeights& pw, float* out_sdf,
                   float* out_feat, int64_t P, int passes, cudaStream_t st, const int* P_dev) {
  return launch_tc(src, packed, pw, out_sdf, out_feat, nullptr, P, passes, -1, nullptr, nullptr, nullptr, st, P_dev);
}

int tc_sdf_forward_grad(const float* x, const void* packed, const PeWeights& pw, float* out_sdf, float* out_feat,
                        float* out_grad, int64_t P, int passes, cudaStream_t st) {
  PointSource src = {};
  src.x = x;
  src.S = 1;
  return launch_tc(src, packed, pw, out_sdf, out_feat, out_grad, P, passes, -1, nullptr, nullptr, nullptr, st);
}

}  // namespace recmv

using namespace recmv;

// ---------------------------------------------------------------------------------------------------------
// Deformer network: MLPTranslator (model/Deformer.py:141-206) + LBSkinner.forward (Deformer.py:406-445)
// ---------------------------------------------------------------------------------------------------------
namespace recmv {
namespace {
constexpr int kDefLayers = 5, kDefPanels = 35;
constexpr int def_in(int l) { return l == 0 ? 167 : 512; }
constexpr int def_out(int l) { return l == 4 ? 3 : 512; }
constexpr int def_pbase(int l) { return l == 0 ? 0 : 3 + 8 * (l - 1); }
constexpr int def_npanels(int l) { return l == 0 ? 3 : 8; }
struct DeformLayout { size_t bias_off, f16_off, total; };
DeformLayout deform_layout() {
  DeformLayout L;
  L.bias_off = 0;
  L.f16_off = ((size_t)kDefLayers * 512 * 4 + 1023) & ~(size_t)1023;
  L.total = L.f16_off + (size_t)2 * kDefPanels * 512 * 64 * 2;
  return L;
}

__global__ void __launch_bounds__(256) pack_plain_layer_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                                               int out, int in, int pbase, int npanels,
                                                               float* __restrict__ bpad, __half* __restrict__ planes,
                                                               int total_panels, DevStatus* status) {
  int64_t total = (int64_t)npanels * 512 * 64;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pi = (int)(i / (512 * 64)), n = (int)((i / 64) % 512), kk = (int)(i % 64);
    const int k = pi * 64 + kk;
    float v = (n < out && k < in) ? W[(size_t)n * in + k] * kWgtScale : 0.f;
    if (!(fabsf(v) < 65504.f)) { report_range(status, 1200); v = fminf(fmaxf(v, -65504.f), 65504.f); }
    const __half h = __float2half_rn(v);
    const size_t o = ((size_t)(pbase + pi) * 512 + n) * 64 + kk;
    planes[o] = h;
    planes[(size_t)total_panels * 512 * 64 + o] = __float2half_rn(v - __half2float(h));
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 512; i += gridDim.x * blockDim.x) bpad[i] = i < out ? b[i] : 0.f;
}
TmapCacheEntry g_def_tmap[16];
}  // namespace
}  // namespace recmv

extern "C" size_t recmv_translator_packed_bytes(void) { return deform_layout().total; }

extern "C" int recmv_translator_pack_weights(const float* W_all, const float* b_all, void* packed,
                                             recmv_stream_t stream) {
  if (!W_all || !b_all || !packed) return RECMV_E_NULL;
  if (((uintptr_t)packed & 1023) != 0) return RECMV_E_SHAPE;
  DeformLayout L = deform_layout();
  char* base = (char*)packed;
  size_t woff = 0, boff = 0;
  void* sd = nullptr;
  int s0 = device_status_record(&sd);
  if (s0) return s0;
  for (int l = 0; l < kDefLayers; ++l) {
    pack_plain_layer_kernel<<<stride_grid((int64_t)def_npanels(l) * 512 * 64, 256, 4), 256, 0, (cudaStream_t)stream>>>(
        W_all + woff, b_all + boff, def_out(l), def_in(l), def_pbase(l), def_npanels(l),
        (float*)(base + L.bias_off) + l * 512, (__half*)(base + L.f16_off), kDefPanels, (DevStatus*)sd);
    int s = launch_status();
    if (s) return s;
    woff += (size_t)def_in(l) * def_out(l);
    boff += def_out(l);
  }
  return RECMV_OK;
}

static int deformer_launch(const float* ps, const float* conds, const int64_t* batch_inds, int64_t points_per_frame,
                           int num_frames, const void* packed, const float* pe_w, const float* A, const float* trans,
                           const recmv_voxel_t* vox, float* out_translated, float* out_offset, float* out_posed,
                           float* out_jac, int64_t P, int mode, recmv_stream_t stream);

extern "C" int recmv_deformer_fwd_jac(const float* ps, const float* conds, const int64_t* batch_inds,
                                      int64_t points_per_frame, int num_frames, const void* packed, const float* pe_w,
                                      const float* A, const float* trans, const recmv_voxel_t* vox, float* out_translated,
                                      float* out_offset, float* out_posed, float* out_jac, int64_t P, int mode,
                                      recmv_stream_t stream) {
  if (!out_jac) return RECMV_E_NULL;
  return deformer_launch(ps, conds, batch_inds, points_per_frame, num_frames, packed, pe_w, A, trans, vox, out_translated,
                         out_offset, out_posed, out_jac, P, mode, stream);
}

extern "C" int recmv_deformer_fwd(const float* ps, const float* conds, const int64_t* batch_inds,
                                  int64_t points_per_frame, int num_frames, const void* packed, const float* pe_w,
                                  const float* A, const float* trans, const recmv_voxel_t* vox, float* out_translated,
                                  float* out_offset, float* out_posed, int64_t P, int mode, recmv_stream_t stream) {
  return deformer_launch(ps, conds, batch_inds, points_per_frame, num_frames, packed, pe_w, A, trans, vox, out_translated,
                         out_offset, out_posed, nullptr, P, mode, stream);
}

static int deformer_launch(const float* ps, const float* conds, const int64_t* batch_inds, int64_t points_per_frame,
                           int num_frames, const void* packed, const float* pe_w, const float* A, const float* trans,
                           const recmv_voxel_t* vox, float* out_translated, float* out_offset, float* out_posed,
                           float* out_jac, int64_t P, int mode, recmv_stream_t stream) {
  if (P < 0 || num_frames <= 0) return RECMV_E_SHAPE;
  if (P == 0) return RECMV_OK;
  if (!ps || !conds || !packed || !pe_w) return RECMV_E_NULL;
  if (out_posed && (!A || !trans || !vox || !vox->ws_cl)) return RECMV_E_NULL;
  if (mode != RECMV_MLP_TC_F16X3 && mode != RECMV_MLP_TC_F16X1) return RECMV_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  cudaGetDevice(&dev);
  DeformLayout L = deform_layout();
  const char* pb = (const char*)packed;
  TmapCacheEntry& tc_ = g_def_tmap[dev & 15];
  if (tc_.base != pb + L.f16_off) {
    int s = make_panel_tmap(&tc_.m128, pb + L.f16_off, (uint64_t)2 * kDefPanels * 512, 128);
    if (s) return s;
    tc_.base = pb + L.f16_off;
  }
  int s0 = tc_prepare_launch(dev);
  if (s0) return s0;
  DevStatus* sd = nullptr;
  s0 = tc_status_record(dev, &sd);
  if (s0) return s0;
  if (sd->code != 0) return RECMV_E_DEVICE;
  TcParams prm = {};
  prm.src.x = ps; prm.src.S = 1;
  for (int i = 0; i < 12; ++i) prm.pw.w[i] = pe_w[i];
  prm.bias = (const float*)(pb + L.bias_off);
  prm.P = P; prm.passes = mode == RECMV_MLP_TC_F16X3 ? 3 : 1; prm.status = sd; prm.dbg_layer = -1;
  set_acc_scales(prm);
  prm.conds = conds; prm.batch_inds = (const long long*)batch_inds; prm.points_per_frame = points_per_frame;
  prm.num_frames = num_frames; prm.bones = A; prm.trans = trans;
  if (vox) prm.vox = to_voxel(vox);
  prm.out_translated = out_translated; prm.out_offset = out_offset; prm.out_posed = out_posed; prm.out_jac = out_jac;
  const int pts_per_tile = out_jac ? 32 : 128;   // forward mode: 4 rows per point
  int64_t tiles = (P + pts_per_tile - 1) / pts_per_tile;
  int64_t want = (tiles + kPairs - 1) / kPairs;
  int maxc = tc_max_clusters(dev);
  int clusters = (int)(want < maxc ? want : maxc);
  if (out_jac)
    sdf_tc_kernel<true, 1><<<clusters * 2 * kPairs, kThreads, kSmemBytes, st>>>(tc_.m128, prm);
  else
    sdf_tc_kernel<false, 1><<<clusters * 2 * kPairs, kThreads, kSmemBytes, st>>>(tc_.m128, prm);
  return launch_status();
}

// ---------------------------------------------------------------------------------------------------------
// Colour network: RenderingNetwork_view_norm, mode 'idr' (model/RenderNet.py:59-96)
// ---------------------------------------------------------------------------------------------------------
namespace recmv {
namespace {
constexpr int kRgbLayers = 5, kRgbPanels = 37;
constexpr int rgb_in(int l) { return l == 0 ? 289 : 512; }
constexpr int rgb_out(int l) { return l == 4 ? 3 : 512; }
constexpr int rgb_pbase(int l) { return l == 0 ? 0 : 5 + 8 * (l - 1); }
constexpr int rgb_npanels(int l) { return l == 0 ? 5 : 8; }
DeformLayout rgb_layout() {
  DeformLayout L;
  L.bias_off = 0;
  L.f16_off = ((size_t)kRgbLayers * 512 * 4 + 1023) & ~(size_t)1023;
  L.total = L.f16_off + (size_t)2 * kRgbPanels * 512 * 64 * 2;
  return L;
}
TmapCacheEntry g_rgb_tmap[16];
}  // namespace
}  // namespace recmv

extern "C" size_t recmv_rendernet_packed_bytes(void) { return rgb_layout().total; }

extern "C" int recmv_rendernet_pack_weights(const float* W_all, const float* b_all, void* packed,
                                            recmv_stream_t stream) {
  if (!W_all || !b_all || !packed) return RECMV_E_NULL;
  if (((uintptr_t)packed & 1023) != 0) return RECMV_E_SHAPE;
  DeformLayout L = rgb_layout();
  char* base = (char*)packed;
  size_t woff = 0, boff = 0;
  void* sd = nullptr;
  int s0 = device_status_record(&sd);
  if (s0) return s0;
  for (int l = 0; l < kRgbLayers; ++l) {
    pack_plain_layer_kernel<<<stride_grid((int64_t)rgb_npanels(l) * 512 * 64, 256, 4), 256, 0, (cudaStream_t)stream>>>(
        W_all + woff, b_all + boff, rgb_out(l), rgb_in(l), rgb_pbase(l), rgb_npanels(l),
        (float*)(base + L.bias_off) + l * 512, (__half*)(base + L.f16_off), kRgbPanels, (DevStatus*)sd);
    int s = launch_status();
    if (s) return s;
    woff += (size_t)rgb_in(l) * rgb_out(l);
    boff += rgb_out(l);
  }
  return RECMV_OK;
}

extern "C" int recmv_rendernet_fwd(const float* points, const float* normals, const float* view_dirs,
                                   const float* feats, const void* packed, const float* pe_w, float* out_rgb,
                                   int64_t P, int mode, recmv_stream_t stream) {
  if (P < 0) return RECMV_E_SHAPE;
  if (P == 0) return RECMV_OK;
  if (!points || !normals || !view_dirs || !feats || !packed || !pe_w || !out_rgb) return RECMV_E_NULL;
  if (mode != RECMV_MLP_TC_F16X3 && mode != RECMV_MLP_TC_F16X1) return RECMV_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  cudaGetDevice(&dev);
  DeformLayout L = rgb_layout();
  const char* pb = (const char*)packed;
  TmapCacheEntry& tc_ = g_rgb_tmap[dev & 15];
  if (tc_.base != pb + L.f16_off) {
    int s = make_panel_tmap(&tc_.m128, pb + L.f16_off, (uint64_t)2 * kRgbPanels * 512, 128);
    if (s) return s;
    tc_.base = pb + L.f16_off;
  }
  int s0 = tc_prepare_launch(dev);
  if (s0) return s0;
  DevStatus* sd = nullptr;
  s0 = tc_status_record(dev, &sd);
  if (s0) return s0;
  if (sd->code != 0) return RECMV_E_DEVICE;
  TcParams prm = {};
  prm.src.x = points; prm.src.S = 1;
  for (int i = 0; i < 12; ++i) prm.pw.w[i] = i < 8 ? pe_w[i] : 0.f;
  prm.bias = (const float*)(pb + L.bias_off);
  prm.P = P; prm.passes = mode == RECMV_MLP_TC_F16X3 ? 3 : 1; prm.status = sd; prm.dbg_layer = -1;
  set_acc_scales(prm);
  prm.normals = normals; prm.view_dirs = view_dirs; prm.feats = feats; prm.out_rgb = out_rgb;
  int64_t tiles = (P + 127) / 128;
  int64_t want = (tiles + kPairs - 1) / kPairs;
  int maxc = tc_max_clusters(dev);
  int clusters = (int)(want < maxc ? want : maxc);
  sdf_tc_kernel<false, 2><<<clusters * 2 * kPairs, kThreads, kSmemBytes, st>>>(tc_.m128, prm);
  return launch_status();
}

// Diagnostics / calibration: relative gain per 64-wide K block applied to the raw accumulators of a precision mode
// (RECMV_MLP_TC_F16X3 or RECMV_MLP_TC_F16X1) to compensate the tensor core's truncating accumulation.
extern "C" int recmv_tc_set_acc_gain(int mode, float gain_per_kblock) {
  if (mode != RECMV_MLP_TC_F16X3 && mode != RECMV_MLP_TC_F16X1) return RECMV_E_DTYPE;
  if (!(fabsf(gain_per_kblock) < 1e-4f)) return RECMV_E_RANGE;
  g_acc_gain_kb[mode == RECMV_MLP_TC_F16X3 ? 0 : 1] = gain_per_kblock;
  return RECMV_OK;
}

// Reports (without synchronising) whether any tcgen05 launch on the current device has aborted on a bounded
// wait since the last check: 0 = none, RECMV_E_DEVICE otherwise; info[3] = {code, barrier tag, block}.
extern "C" int recmv_check_async_errors(int* info, int clear) {
  int dev = 0;
  cudaGetDevice(&dev);
  DevStatus* h = g_status_host[dev & 15];
  if (!h || h->code == 0) {
    if (info) { info[0] = 0; info[1] = 0; info[2] = 0; }
    return RECMV_OK;
  }
  if (info) { info[0] = h->code; info[1] = h->detail; info[2] = h->block; }
  if (clear) { h->code = 0; h->detail = 0; h->block = 0; }
  return RECMV_E_DEVICE;
}

// Diagnostics entry (see include/recmv_b200_diag.h): runs the tcgen05 path on canonical points and returns
// the device status record (which bounded wait timed out, if any) plus the raw accumulator of one layer.
extern "C" int recmv_sdf_mlp_tc_debug(const float* x, const void* packed, const float* pe_w, float* out_sdf,
                                      float* out_feat, int64_t P, int passes, int dbg_layer, float* dbg_out,
                                      int* status_host, unsigned long long* trace, recmv_stream_t stream) {
  if (P <= 0) return RECMV_E_SHAPE;
  if (!x || !packed || !pe_w || !out_sdf || !status_host) return RECMV_E_NULL;
  PointSource src = {};
  src.x = x;
  src.S = 1;
  PeWeights pw;
  for (int i = 0; i < 12; ++i) pw.w[i] = pe_w[i];
  return launch_tc(src, packed, pw, out_sdf, out_feat, nullptr, P, passes, dbg_layer, dbg_out, status_host, trace,
                   (cudaStream_t)stream);
}
