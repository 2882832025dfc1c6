// pinn_tc2_kernels.cuh -- warp-specialised, software-pipelined variant of the tcgen05 kernel (pinn_tc_kernels.cuh).
//
// Same math, same operand formats (bf16x3, X16 shared-memory layout, weight operand in TMEM).  What changes is the
// schedule: a tile is two 16-point sub-tiles A and B that are always one MMA batch apart, so that the tensor core
// works on one sub-tile while the 16 element-wise warps work on the other:
//
//   element-wise warps (0..15)   EW(A,l)  EW(B,l)  EW(A,l+1)  EW(B,l+1) ...      signal rdy[h] after each stage
//   tensor core                       MMA(A,l+1)  MMA(B,l+1)  MMA(A,l+2) ...      commit -> bars[h] (, dwd[h])
//   producer warps (16..19)      restage the weight operand split by split as the previous step's last MMAs release
//                                it, then flush the previous reverse step's dW partial (dwd[1] -> dwfree)
//   issuer warp (20)             per sub-tile: wait rdy[h]; rows MMAs, group by group as the splits become ready
//                                (wr[0..2]); (wait dwfree,) dW MMAs
//
// Weight operand hand-over.  Both sub-tiles use the same layer's operand, so step n+1's operand can only replace
// step n's once B's rows MMAs are through with it -- and A's next rows MMAs (hence the element-wise warps) wait
// for that.  To keep that gap short the three bf16 splits w1, w2, w3 are handed over individually:
//   * the rows MMAs run in three groups  G0 = w2*(a2, a1),  G1 = w3*a1,  G2 = w1*(a3, a2, a1)  (main term last);
//     B's batch commits wf1 after G0 and wf2 after G1, so w2 and w3 are restaged while B's batch is still running;
//   * w1 alternates between two TMEM slots from step to step, so the next w1 is staged a whole step early;
//   * the issuer waits per group, so G0 of the next step runs while w3 is still in flight.
// Four 52-column slots fit next to 2 x D and dW because a split only has 50 live columns: the last k-step of a slot
// reads 4 columns of its right-hand neighbour, which multiply the zero pad rows 100..111 of the activation arrays
// (any finite value will do); the last slot's k-step 6 lives in a 4-column stub in front of slot 0.
//
// The element-wise warps never touch weights, never issue MMAs and never flush dW.
//
// Forward contractions: exact main term (round 2).  The tensor core's fp32 accumulate truncates (scripts/umma_round_probe.py:
// the 16 products and the accumulator are aligned to the largest exponent among them, each addend cut to a 26-bit window,
// the sum cut to 24 bits, all toward zero), which biases every pre-activation toward zero by ~0.5 ulp per full-magnitude
// k-step -- coherently over all points, i.e. like a relative weight perturbation of ~1e-7; near a minimum the gradient
// amplifies that ~1000x (2.5e-5 on the nearly converged Laplace problem; host model: tests/host_harness, mode 4).  So in the
// FORWARD sweep the leading product w1*a1 gets an accumulator of its own in which every sum is EXACT:
//   * w1 = 8 significant bits of w but never finer than 2^-14 of its row's leading bit (pinn_pack_weights_tc; w1 + w2 + w3
//     still equals w bit for bit for every weight within 7 binades of the row maximum),
//   * a1 = a rounded to 7 bits below the leading bit of max |a| over the warp's 32 neurons, per (channel, point) column
//     (store_quad_fwd; a2, a3 = floating bf16 splits of the remainder as before),
//   so all products of a column sit on one grid and span < 24 bits: no truncation can occur.  That accumulator, Dm, is shared
//   by the two sub-tiles and aliased onto the dW columns, which are idle in the forward sweep; the other five products
//   accumulate in D_h as before, and the epilogue adds the two in round-to-nearest fp32.  The main term stays the LAST group of
//   a batch: by then (~0.8 us after the batch started) the element-wise warps have long read the other sub-tile's main term
//   out of Dm, so the wait (dsfree) is a formality and the two sub-tiles stay decoupled.  The reverse sweep keeps the one-
//   accumulator scheme: its truncation is not coherent in the weights and costs nothing measurable (host model mode 0:4).
#pragma once

#include <cstdio>
#include <type_traits>

#include "pinn_tc_kernels.cuh"

namespace pinn {

template <class SIG_, int M_, int H_, bool BWD_>
struct TC2fg {
  using SIG = SIG_;
  static constexpr int M = M_, H = H_;
  static constexpr bool BWD = BWD_;
  static constexpr int C = SIG::C, D = SIG::D;
  static constexpr int NEW = 512;                 // element-wise threads (16 warps)
  static constexpr int NPR = 128;                 // producer threads (4 warps, one per TMEM lane quarter)
  static constexpr int NIS = 32;                  // MMA issuer warp
  static constexpr int NTHALL = NEW + NPR + NIS;
  static constexpr int TS = 16, NSUB = 2, T = TS * NSUB;
  static constexpr int R = C * TS;
  static constexpr int JB = (R / 8) * 128;
  static constexpr int ARR = 14 * JB;
  // Split arrays: a^{l}/a^{l-1} (3 per sub-tile) and, for the reverse sweep, zbar^{l} (3 per sub-tile, or 3 shared by
  // both sub-tiles when two private sets do not fit: C = 6).  With the shared set the element-wise warps of one
  // sub-tile still overlap the other's MMAs with everything except their final zbar stores.
  static constexpr bool SHZ = (C > 5);
  // separate instantiations of the top / bottom reverse stages: measured -14 % (NS2D, m = 3), -1 % (C <= 5, m = 1),
  // +5 % (C = 6, m = 1: more spills)
#ifdef PINN_EXP_PEEL_C5          // what-if builds
  static constexpr bool PEEL = (M > 1) || (C <= 4) || (C == 5 && PINN_EXP_PEEL_C5);
#else
  static constexpr bool PEEL = (M > 1) || (C <= 5);
#endif
  // sub-tile loop of a reverse stage: unrolled for C <= 5, rolled for C = 6 (measured: Poisson2D 11.4 vs 11.8 ms, Burgers2D 16.4 vs 15.7 ms)
#ifdef PINN_EXP_REVUNROLL_C5
  static constexpr int REV_UNROLL = (C <= 4) ? NSUB : (C == 5 ? PINN_EXP_REVUNROLL_C5 : 1);
#else
  static constexpr int REV_UNROLL = (C <= 5) ? NSUB : 1;
#endif
  // Reverse-sweep schedule.  Joint (default): per sub-tile a^{l-1}, zbar^l, then the issuer runs rows(h) + dW(h).
  // Split (what-if, private zbar arrays only): the element-wise warps first turn BOTH sub-tiles' abar^l into zbar^l (the
  // operand of the rows MMAs), then rebuild both a^{l-1} (second operand of the dW contraction); the issuer runs
  // rows(A) rows(B) dW(A) dW(B).  Measured on B200 (1 Mi points, fwd+reverse): Poisson2D 8.82 -> 9.17 ms, Heat2D_Multiscale
  // 10.53 -> 11.51 ms -- the stage no longer waits for the previous dW contraction at its start, but zbar^l(h) cannot be
  // stored before dW(h) of the previous stage has read the arrays, and that contraction is now issued after both rows
  // batches, so the wait moves into the middle of the stage and grows (event trace: profiles/r1_lap_summary.md).
  static constexpr bool SPLIT_REV = false;
  static constexpr int NARR = 3 * NSUB + (BWD ? (SHZ ? 3 : 3 * NSUB) : 0);
  // TMEM columns: D_A | D_B | dW (112) | stub (4) | four weight-split slots at a 52-column stride (the last one 48 wide).
  // The forward sweep's exact main-term accumulator Dm lives in the dW columns (idle then): one per sub-tile where 2R columns
  // fit (DM2: R <= 64, i.e. C <= 4), else one shared by both sub-tiles (R <= 112) and handed over through `dsfree`.
  static constexpr bool DM2 = (2 * R + 2 * R + 4 + 3 * 52 + 48 <= 512);
  static constexpr int tcDmStride = DM2 ? R : 0;
  static constexpr int tcDWWidth = (DM2 && 2 * R > 112) ? 2 * R : 112;
  static constexpr int tcD = 0, tcDstride = R, tcDW = 2 * R, tcStub = 2 * R + tcDWWidth, tcW = tcStub + 4, tcWStride = 52, tcTotal = 512;
  static constexpr int tcWEnd = tcW + 3 * tcWStride + 48;
  static constexpr int oSmall = NARR * ARR;
  static constexpr int oW1 = oSmall;
  static constexpr int oBias = oW1 + H * D * 4;
  static constexpr int oW6 = oBias + kMaxLayers * H * 4;
  static constexpr int oB6 = oW6 + M * H * 4;
  static constexpr int oX = oB6 + 16;
  static constexpr int oAux = oX + T * D * 4;
  static constexpr int oLw = oAux + T * PINN_MAX_AUX * 4;      // per-point direction weights of the Laplacian channel [T][PINN_MAX_DIM]
  static constexpr int oUo = oLw + (SIG::kLapPW ? T * PINN_MAX_DIM * 4 : 0);
  static constexpr int oUb = oUo + NSUB * M * R * 4;
  static constexpr int oRs = oUb + NSUB * M * R * 4;
  static constexpr int oRed = oRs + PINN_MAX_PDE * T * 4;
  static constexpr int oLred = (oRed + 4 * 128 * 4 + 7) / 8 * 8;
  static constexpr int oBar = oLred + PINN_MAX_PDE * T * 8;    // 18 mbarriers + tmem slot
  static constexpr int oKp = oBar + 160;
  static constexpr int nBytes = oKp + (int)((sizeof(KindParams) + 15) / 16 * 16);
  static constexpr size_t smem_bytes = (size_t)nBytes;

  static_assert(H == 100, "tensor-core path is instantiated for H = 100");
  static_assert(R % 16 == 0 && R <= 96, "two sub-tiles in flight need R = C*16 <= 96");
  static_assert(tcWEnd <= tcTotal && (tcWEnd - tcStub) % 16 == 0, "TMEM budget");
  static_assert(oSmall % 16 == 0 && oUo % 16 == 0, "alignment");
  static_assert(nBytes - oSmall >= 2 * JB, "A-operand over-read of the last Zbar array must stay inside the allocation");
  static_assert(smem_bytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// -DPINN_TRACE (debug builds only, see scripts/trace_timeline.py): lane 0 of one element-wise warp, one producer warp
// and the issuer warp of CTA 0 record clock() at the events below for one steady-state tile and print them at exit.
#ifdef PINN_TRACE
#define PINN_TR(ev)                                                                                      \
  do {                                                                                                   \
    if (tr_on && tr_n < 160) tr_buf[tr_n++] = ((unsigned)(ev) << 24) | ((unsigned)clock() & 0xffffffu); \
  } while (0)
#else
#define PINN_TR(ev) \
  do {              \
  } while (0)
#endif

template <class CFG>
__global__ void __launch_bounds__(CFG::NTHALL, 1) pinn_tc2_kernel(const TcLaunchParams p) {
  using SIG = typename CFG::SIG;
  constexpr int C = CFG::C, D = CFG::D, M = CFG::M, H = CFG::H, T = CFG::T, TS = CFG::TS, NSUB = CFG::NSUB, R = CFG::R, JB = CFG::JB,
                ARR = CFG::ARR, NEW = CFG::NEW, NTHALL = CFG::NTHALL;
  constexpr bool BWD = CFG::BWD;

#ifdef PINN_EXP_WARP_ARRIVE       // what-if, measured: one mbarrier arrival per warp instead of per thread -- no gain (see below)
  constexpr int kArr = 32;
#else
  constexpr int kArr = 1;
#endif
  extern __shared__ __align__(1024) unsigned char smraw[];
  float* W1s = reinterpret_cast<float*>(smraw + CFG::oW1);
  float* bs = reinterpret_cast<float*>(smraw + CFG::oBias);
  float* W6s = reinterpret_cast<float*>(smraw + CFG::oW6);
  float* b6s = reinterpret_cast<float*>(smraw + CFG::oB6);
  float* xs = reinterpret_cast<float*>(smraw + CFG::oX);
  float* auxs = reinterpret_cast<float*>(smraw + CFG::oAux);
  float* lws = reinterpret_cast<float*>(smraw + CFG::oLw);
  float* uo = reinterpret_cast<float*>(smraw + CFG::oUo);
  float* ub = reinterpret_cast<float*>(smraw + CFG::oUb);
  float* rs = reinterpret_cast<float*>(smraw + CFG::oRs);
  float* redbuf = reinterpret_cast<float*>(smraw + CFG::oRed);
  double* lred = reinterpret_cast<double*>(smraw + CFG::oLred);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + CFG::oBar);   // [h] MMA batch of sub-tile h complete (tcgen05.commit)
  uint64_t* rdy = bars + 2;                                           // [h] operands of sub-tile h's next batch in place (512 arrivals)
  // wr[0], wr[1]: w1 of an even / odd step staged (the producers may run a step ahead on w1, hence one barrier per
  // slot); wr[2], wr[3]: w2, w3 of this step staged (128 arrivals each)
  uint64_t* wr = bars + 4;
  uint64_t* wfree = bars + 8;                                         // [n&1] rows MMAs of step n complete (w1's slot of step n is free)
  uint64_t* wf1 = bars + 10;                                          // B's rows MMAs are through with w2 (after group G0)
  uint64_t* wf2 = bars + 11;                                          // ... and with w3 (after group G1)
  uint64_t* dwd = bars + 12;                                          // [h] dW MMAs of sub-tile h complete: its arrays may be rewritten
  uint64_t* dwfree = bars + 14;                                       // dW flushed out of TMEM (128 arrivals)
  // rdy2[h]: reverse sweep, split schedule (CFG::SPLIT_REV): the dW operand a^{l-1} of sub-tile h is in place (512 arrivals)
  uint64_t* rdy2 = bars + 15;
  uint64_t* dsfree = bars + 17;                                       // forward: the shared main-term accumulator Dm has been read (16 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  KindParams* kps = reinterpret_cast<KindParams*>(smraw + CFG::oKp);
  // direction weights of the Laplacian channel at tile point `pt`: per point (LapSig<D, true>) or the descriptor's constants
  auto lw_at = [&](int pt) -> const float* {
    if constexpr (SIG::kLapPW) return lws + pt * PINN_MAX_DIM;
    else return kps->lapw;
  };
  // a = 0..2: splits of a^{l} / a^{l-1} of sub-tile h; a = 3..5: splits of zbar^{l}
  auto arr_ptr = [&](int h, int a) -> unsigned char* {
    return smraw + (size_t)(a < 3 ? h * 3 + a : 3 * NSUB + (CFG::SHZ ? 0 : h * 3) + (a - 3)) * ARR;
  };

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3;                        // TMEM lane quarter of this warp
  const int g = (warp >> 2) & 3;                 // element-wise warps: column group (quad of points)
  const int j = 32 * q + lane;                   // neuron (TMEM lane) owned by this thread
#ifdef PINN_EXP_NO_ELEMENTWISE   // what-if timing experiment only
  const bool act = false;
#else
  const bool act = j < H;
#endif
  const int role = warp < NEW / 32 ? 0 : (warp < (NEW + CFG::NPR) / 32 ? 1 : 2);   // 0 element-wise, 1 producer, 2 MMA issuer
  const uint32_t lane_off = (uint32_t)(32 * q) << 16;
  const int pb = 4 * g;
  const int L = p.L;
  const int S = BWD ? p.n_seeds : 0;

  // parameter offsets (nn.Linear order); P < 2^31, 32-bit keeps them out of register pairs
  const int offB1 = H * D;
  const int offHid = offB1 + H;
  const int strideHid = H * H + H;
  const int offW6 = offHid + (L - 1) * strideHid;
  const int offB6 = offW6 + M * H;

  // ---------------- one-time setup (all 20 warps) ----------------
  for (int i = tid; i < H * D; i += NTHALL) W1s[i] = p.params[i];
  for (int i = tid; i < H; i += NTHALL) bs[i] = p.params[offB1 + i];
  for (int l = 2; l <= L; ++l)
    for (int i = tid; i < H; i += NTHALL) bs[(l - 1) * H + i] = p.params[offHid + (l - 2) * strideHid + H * H + i];
  for (int i = tid; i < M * H; i += NTHALL) W6s[i] = p.params[offW6 + i];
  if (tid < M) b6s[tid] = p.params[offB6 + tid];
  for (int i = tid; i < CFG::NARR * ARR / 16; i += NTHALL) reinterpret_cast<uint4*>(smraw)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < PINN_MAX_PDE * T) lred[tid] = 0.0;
  if (tid == 0) {
    mbar_init_n(&bars[0], 1);
    mbar_init_n(&bars[1], 1);
    // one arrival per thread.  -DPINN_EXP_WARP_ARRIVE (every thread fences its own writes, the warp synchronises, lane 0 arrives:
    // 16 / 4 arrivals on a barrier word instead of 512 / 128) was measured at 10.16 ms against 9.84 ms (Poisson2D_Classic, same
    // box; Heat2D_Multiscale unchanged): the arrivals do not queue noticeably, the extra warp synchronisation costs
    mbar_init_n(&rdy[0], NEW / kArr);
    mbar_init_n(&rdy[1], NEW / kArr);
    mbar_init_n(&rdy2[0], NEW / kArr);
    mbar_init_n(&rdy2[1], NEW / kArr);
    for (int i = 0; i < 4; ++i) mbar_init_n(&wr[i], CFG::NPR / kArr);
    mbar_init_n(&wfree[0], 1);
    mbar_init_n(&wfree[1], 1);
    mbar_init_n(wf1, 1);
    mbar_init_n(wf2, 1);
    mbar_init_n(&dwd[0], 1);
    mbar_init_n(&dwd[1], 1);
    mbar_init_n(dwfree, CFG::NPR / kArr);
    mbar_init_n(dsfree, NEW / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *kps = p.kp;
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, CFG::tcTotal);
  float* gp = nullptr;
  if constexpr (BWD) {
    gp = p.gpart + (size_t)blockIdx.x * S * p.Ppad;
    for (long long i = tid; i < (long long)S * p.Ppad; i += NTHALL) gp[i] = 0.f;
    float4* gz = reinterpret_cast<float4*>(p.gwt) + (size_t)blockIdx.x * S * (L - 1) * kTcGwtLayer4;
    for (long long i = tid; i < (long long)S * (L - 1) * kTcGwtLayer4; i += NTHALL) gz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __threadfence();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // shfl: lets ptxas treat TMEM addresses as warp-uniform
  const uint32_t tDW = tbase + CFG::tcDW, tStub = tbase + CFG::tcStub, tW = tbase + CFG::tcW;
  auto tD = [&](int h) -> uint32_t { return tbase + CFG::tcD + h * CFG::tcDstride; };
  const long long ntiles = (p.n_rows + T - 1) / T;
#ifdef PINN_TRACE
  unsigned tr_buf[160];
  int tr_n = 0;
  bool tr_on = false;
  const bool tr_thread = blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 16 || warp == 20);
#endif

  // the (W operand, reverse?) sequence of one tile, identical for the producer warps and the issuer warp:
  //   forward  l = 2..L+1 : W_l (W_{L+1} = zero-padded output layer)
  //   reverse  s, l = L..2: W_l^T, plus the dW_l contraction
  if (role != 0) {
    // =====================================================================================================
    // producer warps: stage weight operands in TMEM, flush dW partials.
    // MMA issuer: its own warp.
    // =====================================================================================================
    const bool do_prod = (role == 1);
    uint32_t wfp[2] = {0, 0}, wf1p = 0, wf2p = 0, dwph = 0;   // producer-side parities of wfree[], wf1, wf2, dwd[1]
    uint4 a[13];
    // slot of split sp at step n: w1 alternates between slot 0 and slot 3, w2 -> slot 1, w3 -> slot 2
    auto slot_of = [&](int sp, long long n) -> int { return sp == 0 ? ((n & 1) ? 3 : 0) : sp; };
    auto load_split = [&](const uint4* Wp, int sp) {
      // one split (52 columns incl. 2 zero pad columns), all 13 loads in flight together: the operand is L2-resident
      // and four warps have to cover the L2 latency with their own memory-level parallelism
#pragma unroll
      for (int i = 0; i < 13; ++i) {
        const int gg = i >> 2, qd = i & 3;
        a[i] = Wp[(size_t)(4 * sp + qd) * 512 + 128 * gg + 32 * q + lane];
      }
    };
    auto store_split = [&](int slot) {
      const uint32_t base = tW + lane_off + (uint32_t)(CFG::tcWStride * slot);
#pragma unroll
      for (int gg = 0; gg < 3; ++gg) {
        uint32_t v[16];
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          v[4 * qd] = a[4 * gg + qd].x; v[4 * qd + 1] = a[4 * gg + qd].y; v[4 * qd + 2] = a[4 * gg + qd].z; v[4 * qd + 3] = a[4 * gg + qd].w;
        }
        tmem_st16(base + 16 * gg, v);
      }
      const uint32_t tail = (slot == 3) ? tStub + lane_off : base + 48;   // columns 48..51 (k-step 6)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tail), "r"(a[12].x), "r"(a[12].y), "r"(a[12].z), "r"(a[12].w)
                   : "memory");
      umma::wait_st();
    };
    auto arrive = [&](uint64_t* bar) {
      umma::fence_before_sync();
      if constexpr (kArr == 32) {
        __syncwarp();
        if (lane != 0) return;
      }
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_addr(bar)) : "memory");
    };
    if (do_prod) {
      // every column a padded k-step may read must hold a finite value from the start
      uint32_t z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll 1
      for (int c = 0; c < CFG::tcWEnd - CFG::tcStub; c += 16) tmem_st16(tStub + lane_off + c, z);
      umma::wait_st();
    }
    // dW_l (TMEM, 100 x 100 used) -> this CTA's private partial [k/4][n][k%4] in L2, read-modify-write through registers (single
    // writer per address, so the partial is deterministic).  Two rounds of 52 / 48 columns; the global loads of a round are all
    // in flight together, the TMEM loads go 4 columns at a time.
    // Tried in round 2, same box, 1 Mi points (profiles/r2_whatif.md): 128-bit `red.global.add` instead of the read-modify-write
    // (-DPINN_FLUSH_RED: half the L2 bytes, no load) -- SLOWER, Poisson2D 9.85 -> 10.32 ms, Heat2D_Multiscale 12.32 -> 12.64 ms
    // (the L2 atomic units serialise the lanes of a vector reduction and hold up the stash / weight traffic behind them);
    // rounds of 32 columns with all TMEM loads of a round in flight (-DPINN_FLUSH_BATCHED: 4 TMEM round trips instead of 25) --
    // SLOWER too, 9.84 -> 10.11 ms and 12.35 -> 12.70 ms: the flush is not latency-bound, and its burst of TMEM reads then
    // collides with the rows MMAs of the next step.
    auto red_add4 = [&](float4* dst, float a, float b, float c, float d) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
    };
    auto flush_dw = [&](int s, int l) {
#ifdef PINN_EXP_NO_FLUSH
      return;
#endif
      float4* gt = reinterpret_cast<float4*>(p.gwt) + ((size_t)(blockIdx.x * S + s) * (L - 1) + (l - 2)) * kTcGwtLayer4 + j;
#ifdef PINN_FLUSH_RED
#pragma unroll
      for (int rd = 0; rd < 3; ++rd) {
        float v0[16], v1[16];
        tmem_ld16(tDW + lane_off + 32 * rd, v0);
        tmem_ld16(tDW + lane_off + 32 * rd + 16, v1);
        umma::wait_ld();
#pragma unroll
        for (int i = 0; i < 4; ++i) red_add4(gt + (size_t)(8 * rd + i) * 128, v0[4 * i], v0[4 * i + 1], v0[4 * i + 2], v0[4 * i + 3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) red_add4(gt + (size_t)(8 * rd + 4 + i) * 128, v1[4 * i], v1[4 * i + 1], v1[4 * i + 2], v1[4 * i + 3]);
      }
      {
        float v3[4];
        umma::tmem_ld4(tDW + lane_off + 96, v3);
        umma::wait_ld();
        red_add4(gt + (size_t)24 * 128, v3[0], v3[1], v3[2], v3[3]);
      }
#elif !defined(PINN_FLUSH_BATCHED)
#pragma unroll
      for (int rd = 0; rd < 2; ++rd) {
        constexpr int NV = 13;
        const int c0 = 52 * rd;
        const int nv = (rd == 0) ? 13 : 12;
        float4 acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (i < nv) acc[i] = gt[(size_t)(c0 / 4 + i) * 128];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (i < nv) {
            float v[4];
            umma::tmem_ld4(tDW + lane_off + c0 + 4 * i, v);
            umma::wait_ld();
            acc[i].x += v[0]; acc[i].y += v[1]; acc[i].z += v[2]; acc[i].w += v[3];
          }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (i < nv) gt[(size_t)(c0 / 4 + i) * 128] = acc[i];
      }
#else
      float4 tail = gt[(size_t)24 * 128];                      // columns 96..99 ride along with the first round
#pragma unroll
      for (int rd = 0; rd < 3; ++rd) {
        float4 acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = gt[(size_t)(8 * rd + i) * 128];
        float v0[16], v1[16];
        tmem_ld16(tDW + lane_off + 32 * rd, v0);
        tmem_ld16(tDW + lane_off + 32 * rd + 16, v1);
        float v3[4];
        if (rd == 0) umma::tmem_ld4(tDW + lane_off + 96, v3);
        umma::wait_ld();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i].x += v0[4 * i]; acc[i].y += v0[4 * i + 1]; acc[i].z += v0[4 * i + 2]; acc[i].w += v0[4 * i + 3];
          acc[4 + i].x += v1[4 * i]; acc[4 + i].y += v1[4 * i + 1]; acc[4 + i].z += v1[4 * i + 2]; acc[4 + i].w += v1[4 * i + 3];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) gt[(size_t)(8 * rd + i) * 128] = acc[i];
        if (rd == 0) {
          tail.x += v3[0]; tail.y += v3[1]; tail.z += v3[2]; tail.w += v3[3];
          gt[(size_t)24 * 128] = tail;
        }
      }
#endif
    };
    long long nstep = 0;                   // pipeline steps so far (forward and reverse, all tiles)
    int pend_s = -1, pend_l = -1;          // dW of the previous reverse step, still in TMEM
    auto producer_stage = [&](const uint4* Wp) {
      const long long n = nstep;
#ifdef PINN_EXP_NO_W
      if (n >= 2) { mbar_wait(&wfree[n & 1], wfp[n & 1]); wfp[n & 1] ^= 1; }
      arrive(&wr[n & 1]);
      if (n >= 1) { mbar_wait(wf1, wf1p); wf1p ^= 1; }
      arrive(&wr[2]);
      if (n >= 1) { mbar_wait(wf2, wf2p); wf2p ^= 1; }
      arrive(&wr[3]);
      return;
#endif
      // w1: its slot was last read by step n-2
      load_split(Wp, 0);
      PINN_TR(40);
      if (n >= 2) {
        mbar_wait(&wfree[n & 1], wfp[n & 1]);
        wfp[n & 1] ^= 1;
        umma::fence_after_sync();
      }
      store_split(slot_of(0, n));
      arrive(&wr[n & 1]);
      // w2: free once group G0 of B's rows MMAs of step n-1 is complete
      load_split(Wp, 1);
      if (n >= 1) {
        mbar_wait(wf1, wf1p);
        wf1p ^= 1;
        umma::fence_after_sync();
      }
      PINN_TR(41);
      store_split(1);
      arrive(&wr[2]);
      // w3: free after group G1
      load_split(Wp, 2);
      if (n >= 1) {
        mbar_wait(wf2, wf2p);
        wf2p ^= 1;
        umma::fence_after_sync();
      }
      store_split(2);
      arrive(&wr[3]);
      PINN_TR(42);
    };
    // the flush owed for the previous reverse step
    auto producer_flush = [&]() {
      mbar_wait(&dwd[1], dwph);            // dW of the previous step is complete
      dwph ^= 1;
      umma::fence_after_sync();
      PINN_TR(43);
      flush_dw(pend_s, pend_l);
      PINN_TR(44);
      arrive(dwfree);
      pend_l = -1;
    };
    uint32_t rph[2] = {0, 0}, wrp[4] = {0, 0, 0, 0}, fph = 0;
    bool flush_pending = false;            // a reverse step's dW is still in TMEM (its flush has not been waited for)
    // rows MMAs of sub-tile h: D_h (+)= W[TMEM] * B[smem, MN-major], 7 k-steps x 6 split products in three groups.
    // first: this is the step's first batch (wait for each split as its group comes up); last: it is the step's last
    // batch (release w2 / w3 / everything as the groups retire).
    uint32_t dsph = 0;
    bool dm_used = false;                  // a forward batch has written Dm: the next one may only once that has been read
    auto issue_rows = [&](int h, unsigned char* bbase, uint32_t el, bool first, bool last, auto fwd_c) {
      constexpr bool fwd = decltype(fwd_c)::value;          // compile-time: every operand address below folds to a constant offset
      constexpr uint32_t idesc = umma::make_idesc_bf16(kTcRows, R, 0, 1);
      // G0 = w2*(a2, a1) | G1 = w3*a1 | G2 = w1*(a3, a2, a1), the main term last.  Reverse: all into D_h.  Forward: the five
      // small products into D_h, the main term w1*a1 into Dm (exact, see the header) -- Dm is shared by the two sub-tiles, but
      // by the time a batch reaches its last group the element-wise warps have long read the other sub-tile's main term.
      constexpr int wa[6] = {1, 1, 2, 0, 0, 0}, ab[6] = {1, 0, 0, 2, 1, 0};
      const uint32_t d = tD(h);
      const uint32_t s0 = (uint32_t)(CFG::tcWStride * slot_of(0, nstep));
      const uint32_t s0last = (nstep & 1) ? tStub : tW + s0 + 48;     // k-step 6 of w1 (slot 3 keeps it in the stub)
      const uint32_t lo0 = umma::smem_desc_lo(umma::smem_addr(bbase), JB);
      constexpr uint32_t hi = umma::smem_desc_hi(128);
      uint32_t accum = 0;
#pragma unroll
      for (int t = 0; t < 6; ++t) {
        if (first && (t == 0 || t == 2 || t == 3)) {
          const int wb = (wa[t] == 0) ? (int)(nstep & 1) : wa[t] + 1;
          mbar_wait(&wr[wb], wrp[wb]);
          wrp[wb] ^= 1;
          umma::fence_after_sync();
        }
#ifndef PINN_EXP_NO_DM
        if (fwd && t == 5) {
          // Dm: the element-wise warps must have read the previous forward batch's main term out of it, and the last
          // reverse step's dW (same columns) must have been flushed
          PINN_TR(70 + h);
#ifndef PINN_EXP_DM_NOWAIT
          if constexpr (!CFG::DM2) {
            if (dm_used) {
              mbar_wait(dsfree, dsph);
              dsph ^= 1;
            }
          }
#endif
          PINN_TR(74 + h);
          dm_used = true;
          if (flush_pending) {
            mbar_wait(dwfree, fph);
            fph ^= 1;
            flush_pending = false;
          }
#ifndef PINN_EXP_DM_NOFENCE
          umma::fence_after_sync();
#endif
          PINN_TR(72 + h);
          accum = 0;
        }
#endif
#ifdef PINN_EXP_NO_DM             // what-if timing experiment only: one accumulator as in round 1
        const uint32_t dst = d;
#else
        const uint32_t dst = (fwd && t == 5) ? tDW + (uint32_t)(h * CFG::tcDmStride) : d;
#endif
#pragma unroll
        for (int ks = 0; ks < kTcHP / 16; ++ks) {
          const uint32_t wad = (wa[t] == 0) ? (ks == 6 ? s0last : tW + s0 + (uint32_t)(ks * 8)) : tW + (uint32_t)(wa[t] * CFG::tcWStride + ks * 8);
          umma::mma_bf16_ts_e2(dst, wad, lo0 + (uint32_t)((ab[t] * ARR + ks * 2 * JB) >> 4), hi, idesc, accum, el);
          accum = 1;
        }
        if (last && t == 1) umma::commit_e(wf1, el);
        if (last && t == 2) umma::commit_e(wf2, el);
      }
    };
    auto issue_dw = [&](int h, uint32_t accum, uint32_t el) {
      constexpr uint32_t idesc = umma::make_idesc_bf16(kTcRows, kTcHP, 0, 0);
      const int za[6] = {0, 2, 1, 0, 1, 0};
      const int ab[6] = {2, 0, 1, 1, 0, 0};
      const uint32_t zlo0 = umma::smem_desc_lo(umma::smem_addr(arr_ptr(h, 3)), 128);
      const uint32_t alo0 = umma::smem_desc_lo(umma::smem_addr(arr_ptr(h, 0)), 128);
      constexpr uint32_t hi = umma::smem_desc_hi(JB);
#pragma unroll
      for (int t = 0; t < 6; ++t) {
#pragma unroll
        for (int ks = 0; ks < R / 16; ++ks) {
          umma::mma_bf16_ss_e2(tDW, zlo0 + (uint32_t)((za[t] * ARR + ks * 256) >> 4), hi, alo0 + (uint32_t)((ab[t] * ARR + ks * 256) >> 4), hi, idesc,
                               accum, el);
          accum = 1;
        }
      }
    };
    // One pipeline step = the batches of both sub-tiles for one layer.  Within a sub-tile the rows MMAs (the only
    // readers of the weight operand, the producers of D) go before the dW MMAs.
    auto issue_sub = [&](int h, bool reverse) {
      PINN_TR(60 + h);
      mbar_wait(&rdy[h], rph[h]);
      rph[h] ^= 1;
      umma::fence_after_sync();
      PINN_TR(62 + h);
      const uint32_t el = umma::elect_one();
      if (reverse) issue_rows(h, arr_ptr(h, 3), el, h == 0, h == NSUB - 1, std::false_type{});
      else issue_rows(h, arr_ptr(h, 0), el, h == 0, h == NSUB - 1, std::true_type{});
      umma::commit_e(&bars[h], el);
      if (h == NSUB - 1) umma::commit_e(&wfree[nstep & 1], el);
      PINN_TR(64 + h);
      if (reverse) {
        if (h == 0 && flush_pending) {
          mbar_wait(dwfree, fph);          // the previous reverse step's dW has left TMEM
          fph ^= 1;
          flush_pending = false;
          umma::fence_after_sync();
        }
        PINN_TR(66 + h);
        issue_dw(h, h == 0 ? 0u : 1u, el);
        umma::commit_e(&dwd[h], el);
        PINN_TR(68 + h);
      }
    };
    // split reverse schedule (CFG::SPLIT_REV): rows(A) rows(B) dW(A) dW(B)
    uint32_t r2ph[2] = {0, 0};
    auto issue_rows_only = [&](int h) {
      PINN_TR(60 + h);
      mbar_wait(&rdy[h], rph[h]);
      rph[h] ^= 1;
      umma::fence_after_sync();
      PINN_TR(62 + h);
      const uint32_t el = umma::elect_one();
      issue_rows(h, arr_ptr(h, 3), el, h == 0, h == NSUB - 1, std::false_type{});
      umma::commit_e(&bars[h], el);
      if (h == NSUB - 1) umma::commit_e(&wfree[nstep & 1], el);
      PINN_TR(64 + h);
    };
    auto issue_dw_only = [&](int h) {
      if (h == 0 && flush_pending) {
        mbar_wait(dwfree, fph);            // the previous reverse step's dW has left TMEM
        fph ^= 1;
        flush_pending = false;
        umma::fence_after_sync();
      }
      mbar_wait(&rdy2[h], r2ph[h]);
      r2ph[h] ^= 1;
      umma::fence_after_sync();
      PINN_TR(66 + h);
      const uint32_t el = umma::elect_one();
      issue_dw(h, h == 0 ? 0u : 1u, el);
      umma::commit_e(&dwd[h], el);
      PINN_TR(68 + h);
    };
    auto step = [&](const uint4* Wp, bool reverse, int s, int l) {
      if (do_prod) {
        producer_stage(Wp);
        if (pend_l >= 2) producer_flush();
        if (reverse) {
          pend_s = s;
          pend_l = l;
          // The tile's last reverse step: flush its dW right away instead of one step late -- the next step is a forward
          // one, whose small-term accumulator lives in the dW columns, and the tensor core idles at the tile boundary anyway.
          if (l == 2 && s == S - 1) producer_flush();
        }
      } else {
        PINN_TR(58);
        if (reverse && CFG::SPLIT_REV) {
          issue_rows_only(0);
          issue_rows_only(1);
          issue_dw_only(0);
          issue_dw_only(1);
        } else {
          issue_sub(0, reverse);
          issue_sub(1, reverse);
        }
        if (reverse) flush_pending = true;
      }
      ++nstep;
    };
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#ifdef PINN_TRACE
      tr_on = tr_thread && (tile / gridDim.x == 3);
#endif
      for (int l = 2; l <= L + 1; ++l) step(l <= L ? p.wn + (size_t)(l - 2) * kTcWMat : p.w6p, false, 0, 0);
      if constexpr (BWD) {
        for (int s = 0; s < S; ++s)
          for (int l = L; l >= 2; --l) step(p.wt + (size_t)(l - 2) * kTcWMat, true, s, l);
      }
    }
    if (do_prod && pend_l >= 2) producer_flush();
  } else {
    // =====================================================================================================
    // element-wise warps
    // =====================================================================================================
    float4* stash = reinterpret_cast<float4*>(p.stash) + (size_t)blockIdx.x * L * NSUB * C * NEW;
    auto stash_at = [&](int l, int h, int c) -> float4* { return stash + ((size_t)((l - 1) * NSUB + h) * C + c) * NEW + tid; };
    const bool stash_discard = (reinterpret_cast<uintptr_t>(p.stash) & 127) == 0;
    // Stash stores / loads carry an L2 evict_last policy: the scratch lines are the last candidates for replacement (and
    // write-back) between their write in the forward sweep and their last read in the reverse sweep, after which they are
    // discarded.  Measured (ncu, Poisson2D_Classic, 1 Mi points): DRAM writes 3.25 GB -> 0.56 GB per launch, same duration.
    // -DPINN_NO_STASH_EVICT_LAST restores plain accesses.
#ifndef PINN_NO_STASH_EVICT_LAST
    uint64_t stash_pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(stash_pol));
    auto stash_st = [&](float4* ptr, float a, float b, float c, float d) {
      asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(a), "f"(b), "f"(c), "f"(d), "l"(stash_pol) : "memory");
    };
    auto stash_ld = [&](const float4* ptr) -> float4 {
      float4 v;
      asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(stash_pol) : "memory");
      return v;
    };
#else
    auto stash_st = [&](float4* ptr, float a, float b, float c, float d) { *ptr = make_float4(a, b, c, d); };
    auto stash_ld = [&](const float4* ptr) -> float4 { return *ptr; };
#endif
    uint32_t ph[2] = {0, 0}, dph[2] = {0, 0};
    bool dwp[2] = {false, false};          // a dW contraction that reads sub-tile h's arrays may still be running

    auto store_quad = [&](unsigned char* base, int r0, float v0, float v1, float v2, float v3) {
      uint32_t a1, a2, a3, b1, b2, b3;
      split3_pair(v0, v1, a1, a2, a3);
      split3_pair(v2, v3, b1, b2, b3);
      const int off = (j >> 3) * JB + (r0 >> 3) * 128 + (j & 7) * 16 + (r0 & 7) * 2;
      *reinterpret_cast<uint2*>(base + off) = make_uint2(a1, b1);
      *reinterpret_cast<uint2*>(base + ARR + off) = make_uint2(a2, b2);
      *reinterpret_cast<uint2*>(base + 2 * ARR + off) = make_uint2(a3, b3);
    };
    // Forward operand of the next layer's rows contraction (see the header): a1 on the column's grid, a2 + a3 the remainder.
    // `mg` = 1.5 * 2^23 * grid per point: (v + mg) - mg rounds v to a multiple of the grid (|v| < 2^22 grids always holds).
    const unsigned wmask = (q == 3) ? ((1u << (H - 96)) - 1u) : 0xffffffffu;     // lanes of this warp that own a neuron
    auto grid_magic = [&](float v0, float v1, float v2, float v3) -> float {
      // leading bit of max |v| over the warp's neurons and the quad's four points -> grid 2^(e - 7) -> magic constant
      // 1.5 * 2^(e + 16).  One grid per (channel, quad, warp): a point whose values are smaller than the quad's largest keeps
      // fewer bits in a1 and more in a2 + a3 -- the sums stay exact, which is all the main term needs.
      const unsigned b0 = __float_as_uint(v0) & 0x7fffffffu, b1 = __float_as_uint(v1) & 0x7fffffffu;
      const unsigned b2 = __float_as_uint(v2) & 0x7fffffffu, b3 = __float_as_uint(v3) & 0x7fffffffu;
      const unsigned mx = __reduce_max_sync(wmask, max(max(b0, b1), max(b2, b3)));
      unsigned e = mx >> 23;
      e = e > 230u ? 230u : e;
      return __uint_as_float(((e + 17u) << 23) | 0x00400000u);
    };
    auto store_quad_fwd = [&](unsigned char* base, int c, int r0, float v0, float v1, float v2, float v3) {
#ifdef PINN_EXP_STATIC_GRID      // what-if timing experiment only (results wrong)
      const float m0 = 1.5f * 65536.0f;
#else
      const float m0 = (c == 0) ? 1.5f * 65536.0f : grid_magic(v0, v1, v2, v3);          // tanh values: |v| < 1, grid 2^-7
#endif
      const float m1 = m0, m2 = m0, m3 = m0;
      const float h0 = __fsub_rn(__fadd_rn(v0, m0), m0), h1 = __fsub_rn(__fadd_rn(v1, m1), m1);
      const float h2 = __fsub_rn(__fadd_rn(v2, m2), m2), h3 = __fsub_rn(__fadd_rn(v3, m3), m3);
      const uint32_t a1 = pack_bf16x2(h0, h1), b1 = pack_bf16x2(h2, h3);        // exact: at most 8 significant bits
      uint32_t a2, a3, b2, b3;
      split2_pair(v0 - h0, v1 - h1, a2, a3);
      split2_pair(v2 - h2, v3 - h3, b2, b3);
      const int off = (j >> 3) * JB + (r0 >> 3) * 128 + (j & 7) * 16 + (r0 & 7) * 2;
      *reinterpret_cast<uint2*>(base + off) = make_uint2(a1, b1);
      *reinterpret_cast<uint2*>(base + ARR + off) = make_uint2(a2, b2);
      *reinterpret_cast<uint2*>(base + 2 * ARR + off) = make_uint2(a3, b3);
    };
    // forward pre-activations: small terms (D_h) + exact main term (Dm), added in round-to-nearest; releases Dm
    auto load_fwd = [&](int h, float (&zq)[C][4]) {
#pragma unroll
      for (int c = 0; c < C; ++c) umma::tmem_ld4(tD(h) + lane_off + c * TS + pb, zq[c]);
#ifdef PINN_EXP_NO_DM
      umma::wait_ld();
      return;
#endif
      const uint32_t tDm = tDW + (uint32_t)(h * CFG::tcDmStride) + lane_off + pb;
      // the main term in two rounds: keeps the live registers of this stage at 1.5 accumulator sets instead of 2
      constexpr int C0 = (C + 1) / 2;
      {
        float sm[C0][4];
#pragma unroll
        for (int c = 0; c < C0; ++c) umma::tmem_ld4(tDm + c * TS, sm[c]);
        umma::wait_ld();
#pragma unroll
        for (int c = 0; c < C0; ++c)
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) zq[c][pp] += sm[c][pp];
      }
      {
        float sm[C - C0][4];
#pragma unroll
        for (int c = C0; c < C; ++c) umma::tmem_ld4(tDm + c * TS, sm[c - C0]);
        umma::wait_ld();
        if constexpr (!CFG::DM2) {
          umma::fence_before_sync();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_addr(dsfree)) : "memory");
        }
        PINN_TR(16 + h);
#pragma unroll
        for (int c = C0; c < C; ++c)
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) zq[c][pp] += sm[c - C0][pp];
      }
    };
    auto signal_ready = [&](int h) {
      umma::fence_async_smem();
      umma::fence_before_sync();
      if constexpr (kArr == 32) {
        __syncwarp();
        if (lane != 0) return;
      }
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_addr(&rdy[h])) : "memory");
    };
    auto signal_ready2 = [&](int h) {
      umma::fence_async_smem();
      umma::fence_before_sync();
      if constexpr (kArr == 32) {
        __syncwarp();
        if (lane != 0) return;
      }
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_addr(&rdy2[h])) : "memory");
    };
    auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NEW) : "memory"); };
    auto wait_bar = [&](int h) {
      mbar_wait(&bars[h], ph[h]);
      ph[h] ^= 1;
      umma::fence_after_sync();
    };
    // before rewriting sub-tile h's arrays: the last dW contraction that reads them must be complete
    auto guard_arrays = [&](int h) {
      if (dwp[h]) {
        mbar_wait(&dwd[h], dph[h]);
        dph[h] ^= 1;
        dwp[h] = false;
      }
    };
    auto reduce_groups_red = [&](float v, float* dst) {
      redbuf[g * 128 + j] = v;
      epi_sync();
      if (g == 0 && act) atomicAdd(dst, (redbuf[j] + redbuf[128 + j]) + (redbuf[256 + j] + redbuf[384 + j]));
      epi_sync();
    };
    auto epi_layer1 = [&](int h) {
      guard_arrays(h);
      if (!act) return;
      float w[D];
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = W1s[j * D + i];
      const float b = bs[j];
      float a[C][4];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        float z0 = b;
#pragma unroll
        for (int i = 0; i < D; ++i) z0 = fmaf(w[i], xs[(h * TS + pb + pp) * D + i], z0);
#pragma unroll
        for (int c = 0; c < C; ++c) a[c][pp] = 0.f;
        a[0][pp] = z0;
        static_for<D>([&](auto J) {
          constexpr int jj = decltype(J)::value;
          if constexpr (SIG::order(jj) >= 1) a[SIG::base(jj)][pp] = w[jj];
        });
      }
      // The stash keeps t0 = tanh(z_0) in channel 0 instead of z_0: nothing downstream needs z_0 itself (tanh_jet_*_t0),
      // and the reverse sweep then recomputes a^{l-1} and the adjoint without a transcendental.
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) a[0][pp] = tanh_tcx<SIG::kAccTanh>(a[0][pp]);
      if constexpr (BWD) {
#pragma unroll
        for (int c = 0; c < C; ++c) stash_st(stash_at(1, h, c), a[c][0], a[c][1], a[c][2], a[c][3]);
      }
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        float z[C], av[C];
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = a[c][pp];
        tanh_jet_fwd_t0<SIG, true>(z, z[0], av, lw_at(h * TS + pb + pp));
#pragma unroll
        for (int c = 0; c < C; ++c) a[c][pp] = av[c];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) store_quad_fwd(arr_ptr(h, 0), c, c * TS + pb, a[c][0], a[c][1], a[c][2], a[c][3]);
    };
    auto epi_hidden = [&](int h, int l) {
      float zq[C][4];
      load_fwd(h, zq);
      if (!act) return;
      const float b = bs[(l - 1) * H + j];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) zq[0][pp] = tanh_tcx<SIG::kAccTanh>(zq[0][pp] + b);      // channel 0 holds t0 from here on (see epi_layer1)
      if constexpr (BWD) {
#pragma unroll
        for (int c = 0; c < C; ++c) stash_st(stash_at(l, h, c), zq[c][0], zq[c][1], zq[c][2], zq[c][3]);
      }
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        float z[C], av[C];
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = zq[c][pp];
        tanh_jet_fwd_t0<SIG>(z, z[0], av, lw_at(h * TS + pb + pp));
#pragma unroll
        for (int c = 0; c < C; ++c) zq[c][pp] = av[c];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) store_quad_fwd(arr_ptr(h, 0), c, c * TS + pb, zq[c][0], zq[c][1], zq[c][2], zq[c][3]);
    };
    auto epi_output = [&](int h) {
      float uq[C][4];
      load_fwd(h, uq);
      if (j < M) {
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) uo[(h * M + j) * R + c * TS + pb + pp] = uq[c][pp] + (c == 0 ? b6s[j] : 0.f);
      }
    };

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#ifdef PINN_TRACE
      tr_on = tr_thread && (tile / gridDim.x == 3);
#endif
      const long long row0 = tile * T;
      for (int i = tid; i < T * D; i += NEW) {
        const long long r = row0 + i / D;
        xs[i] = (r < p.n_rows) ? p.x[row0 * D + i] : 0.f;
      }
      if (p.n_aux > 0) {
        for (int i = tid; i < T * p.n_aux; i += NEW) {
          const int pt = i / p.n_aux, a = i % p.n_aux;
          const long long r = row0 + pt;
          auxs[pt * PINN_MAX_AUX + a] = (r < p.n_rows) ? p.aux[r * p.n_aux + a] : 0.f;
        }
      }
      if constexpr (SIG::kLapPW) {
        epi_sync();                        // the weights read this tile's aux columns
        if (tid < T) lap_point_weights(*kps, auxs + tid * PINN_MAX_AUX, D, lws + tid * PINN_MAX_DIM);
      }
      epi_sync();

      // ================= forward =================
#pragma unroll
      for (int h = 0; h < NSUB; ++h) {
        PINN_TR(1 + h);
        epi_layer1(h);
        signal_ready(h);
        PINN_TR(3 + h);
      }
      for (int l = 2; l <= L + 1; ++l) {
#pragma unroll
        for (int h = 0; h < NSUB; ++h) {
          PINN_TR(10 + h);
          wait_bar(h);
          PINN_TR(12 + h);
          if (l <= L) {
            epi_hidden(h, l);
            signal_ready(h);
          } else {
            epi_output(h);
          }
          PINN_TR(14 + h);
        }
      }
      epi_sync();

      // ================= residual + loss =================
      if (tid < T) {
        const int h = tid / TS, pt = tid % TS;
        float U[M][C];
#pragma unroll
        for (int o = 0; o < M; ++o)
#pragma unroll
          for (int c = 0; c < C; ++c) U[o][c] = uo[(h * M + o) * R + c * TS + pt];
        float r[PINN_MAX_PDE];
        Residual<SIG, M>::eval(*kps, auxs + tid * PINN_MAX_AUX, U, r);
        const bool valid = row0 + tid < p.n_rows;
#pragma unroll
        for (int k = 0; k < PINN_MAX_PDE; ++k) {
          const float rk = valid ? r[k] : 0.f;
          rs[k * T + tid] = rk;
          lred[k * T + tid] += (double)rk * (double)rk;   // per-point running sums live in shared memory, not in registers
        }
        if (p.resid_out != nullptr && valid)
          for (int k = 0; k < p.n_pde; ++k) p.resid_out[(row0 + tid) * p.n_pde + k] = r[k];
      }

      if constexpr (BWD) {
        for (int s = 0; s < S; ++s) {
          float* gs = gp + (size_t)s * p.Ppad;
          if (tid < T) {
            const int h = tid / TS, pt = tid % TS;
            float U[M][C], Ub[M][C];
#pragma unroll
            for (int o = 0; o < M; ++o)
#pragma unroll
              for (int c = 0; c < C; ++c) U[o][c] = uo[(h * M + o) * R + c * TS + pt];
            float rb[PINN_MAX_PDE];
#pragma unroll
            for (int k = 0; k < PINN_MAX_PDE; ++k) rb[k] = (k < p.n_pde) ? p.seeds[s * p.n_pde + k] * (kps->lin_seed ? (row0 + tid < p.n_rows ? 1.0f : 0.f) : 2.0f * rs[k * T + tid] * p.inv_n) : 0.f;
            Residual<SIG, M>::adjoint(*kps, auxs + tid * PINN_MAX_AUX, U, rb, Ub);
#pragma unroll
            for (int o = 0; o < M; ++o)
#pragma unroll
              for (int c = 0; c < C; ++c) ub[(h * M + o) * R + c * TS + pt] = Ub[o][c];
          }
          epi_sync();
          if (tid < M) {
            float sacc = 0.f;
            for (int h = 0; h < NSUB; ++h)
              for (int pq = 0; pq < TS; ++pq) sacc += ub[(h * M + tid) * R + pq];
            atomicAdd(gs + offB6 + tid, sacc);
          }

          // One reverse stage (both sub-tiles) of layer l.  The top (l == L: abar from the output adjoint, dW6) and bottom
          // (l == 1: dW1, no MMAs) stages are separate instantiations where that helps register allocation (CFG::PEEL:
          // their extra live state -- dw6, dw1, the W6 row -- then does not weigh on the L-2 stages in between).
          auto rev_stage = [&](auto top_c, auto bottom_c, const int l) {
            // 0 / 1: known at compile time; 2: decided at run time (one shared instantiation)
            const bool IS_TOP = decltype(top_c)::value == 2 ? (l == L) : (decltype(top_c)::value == 1);
            const bool IS_BOTTOM = decltype(bottom_c)::value == 2 ? (l == 1) : (decltype(bottom_c)::value == 1);
            float dbias = 0.f;
            float dw6[M];
#pragma unroll
            for (int o = 0; o < M; ++o) dw6[o] = 0.f;
            float dw1[D];
#pragma unroll
            for (int i = 0; i < D; ++i) dw1[i] = 0.f;
            auto part1 = [&](int h) {
              // ---- part 1, independent of the tensor core: splits of a^{l-1} (operand of the dW contraction)
              PINN_TR(20 + h);
              if (!IS_BOTTOM) {
                guard_arrays(h);
                PINN_TR(22 + h);
                if (act) {
                  float zq[C][4];
#pragma unroll
                  for (int c = 0; c < C; ++c) {
                    const float4 v = stash_ld(stash_at(l - 1, h, c));
                    zq[c][0] = v.x; zq[c][1] = v.y; zq[c][2] = v.z; zq[c][3] = v.w;
                  }
#pragma unroll
                  for (int pp = 0; pp < 4; ++pp) {
                    float z[C], av[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) z[c] = zq[c][pp];
                    tanh_jet_fwd_t0<SIG>(z, z[0], av, lw_at(h * TS + pb + pp));       // z[0] is the stashed t0
#pragma unroll
                    for (int c = 0; c < C; ++c) zq[c][pp] = av[c];
                  }
#pragma unroll
                  for (int c = 0; c < C; ++c) store_quad(arr_ptr(h, 0), c * TS + pb, zq[c][0], zq[c][1], zq[c][2], zq[c][3]);
                }
              }
            };
            auto part2 = [&](int h) {
              // ---- part 2: abar^l (from D_h, written by the rows MMAs of layer l+1) -> zbar^l
              PINN_TR(24 + h);
              if (!IS_TOP) wait_bar(h);
              PINN_TR(26 + h);
              float ab[C][4];
              if (!IS_TOP) {
#pragma unroll
                for (int c = 0; c < C; ++c) umma::tmem_ld4(tD(h) + lane_off + c * TS + pb, ab[c]);
                umma::wait_ld();
              }
              if (act) {
                float zq[C][4];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                  const float4 v = stash_ld(stash_at(l, h, c));
                  zq[c][0] = v.x; zq[c][1] = v.y; zq[c][2] = v.z; zq[c][3] = v.w;
                }
                if (IS_TOP) {
#pragma unroll
                  for (int pp = 0; pp < 4; ++pp) {
                    float z[C], av[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) z[c] = zq[c][pp];
                    tanh_jet_fwd_t0<SIG>(z, z[0], av, lw_at(h * TS + pb + pp));       // z[0] is the stashed t0
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                      float v = 0.f;
#pragma unroll
                      for (int o = 0; o < M; ++o) {
                        const float u = ub[(h * M + o) * R + c * TS + pb + pp];
                        v = fmaf(u, W6s[o * H + j], v);
                        dw6[o] = fmaf(u, av[c], dw6[o]);
                      }
                      ab[c][pp] = v;
                    }
                  }
                }
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                  float z[C], a_[C], zb[C];
#pragma unroll
                  for (int c = 0; c < C; ++c) {
                    z[c] = zq[c][pp];
                    a_[c] = ab[c][pp];
                    zb[c] = 0.f;
                  }
                  tanh_jet_bwd_t0<SIG>(z, z[0], a_, zb, lw_at(h * TS + pb + pp));  // z[0] is the stashed t0
                  dbias += zb[0];
                  if (IS_BOTTOM) {
                    static_for<D>([&](auto J) {
                      constexpr int jj = decltype(J)::value;
                      dw1[jj] = fmaf(zb[0], xs[(h * TS + pb + pp) * D + jj], dw1[jj]);
                      if constexpr (SIG::order(jj) >= 1) dw1[jj] += zb[SIG::base(jj)];
                    });
                  }
#pragma unroll
                  for (int c = 0; c < C; ++c) ab[c][pp] = zb[c];
                }
#ifndef PINN_NO_STASH_DISCARD
                // That was the last read of z^l of this tile: drop its (dirty) L2 lines instead of letting L2 write them back
                // to HBM -- the stash is scratch, rewritten by the next tile.  One 128-byte line per 8 consecutive threads;
                // the loaded values have been consumed above, so every lane's read of the line has completed.
                if (stash_discard && s == S - 1 && (lane & 7) == 0) {     // (the other seeds' reverse sweeps still read it)
#pragma unroll
                  for (int c = 0; c < C; ++c) asm volatile("discard.global.L2 [%0], 128;" ::"l"(stash_at(l, h, c)) : "memory");
                }
#endif
              }
              if (!IS_BOTTOM) {
                if constexpr (CFG::SHZ) {
                  // shared zbar arrays: the other sub-tile's MMAs (its pending dwd phase) must be done reading them
                  if (dwp[1 - h]) mbar_wait(&dwd[1 - h], dph[1 - h]);
                }
                // split schedule: zbar(h) was last read by the previous stage's dW(h) contraction (so were the a arrays part1
                // rewrites afterwards); in the joint schedule part1 has waited for it already
                if constexpr (CFG::SPLIT_REV) guard_arrays(h);
                if (act) {
#pragma unroll
                  for (int c = 0; c < C; ++c) store_quad(arr_ptr(h, 3), c * TS + pb, ab[c][0], ab[c][1], ab[c][2], ab[c][3]);
                }
              }
              PINN_TR(28 + h);
            };
            if constexpr (CFG::SPLIT_REV) {
#pragma unroll
              for (int h = 0; h < NSUB; ++h) {
                part2(h);
                if (!IS_BOTTOM) signal_ready(h);
              }
              if (!IS_BOTTOM) {
#pragma unroll
                for (int h = 0; h < NSUB; ++h) {
                  part1(h);
                  signal_ready2(h);
                  dwp[h] = true;
                }
              }
            } else {
#pragma unroll(CFG::REV_UNROLL)
              for (int h = 0; h < NSUB; ++h) {
                part1(h);
                part2(h);
                if (!IS_BOTTOM) {
                  signal_ready(h);
                  dwp[h] = true;
                }
              }
            }
            float* gb = gs + (IS_BOTTOM ? offB1 : offHid + (l - 2) * strideHid + H * H);
            reduce_groups_red(dbias, gb + j);
            if (IS_TOP) {
#pragma unroll
              for (int o = 0; o < M; ++o) reduce_groups_red(dw6[o], gs + offW6 + o * H + j);
            }
            if (IS_BOTTOM) {
#pragma unroll
              for (int i = 0; i < D; ++i) reduce_groups_red(dw1[i], gs + j * D + i);
            }
          };
          using Yes = std::integral_constant<int, 1>;
          using No = std::integral_constant<int, 0>;
          using Rt = std::integral_constant<int, 2>;
          if constexpr (CFG::PEEL) {
            rev_stage(Yes{}, No{}, L);
            for (int l = L - 1; l >= 2; --l) rev_stage(No{}, No{}, l);
            rev_stage(No{}, Yes{}, 1);
          } else {
            for (int l = L; l >= 1; --l) rev_stage(Rt{}, Rt{}, l);
          }
        }
      }
      epi_sync();
    }
  }

#ifdef PINN_TRACE
  if (tr_thread)
    for (int i = 0; i < tr_n; ++i) printf("TR %d %u %u\n", warp, tr_buf[i] >> 24, tr_buf[i] & 0xffffffu);
#endif
  // ---------------- per-CTA loss partials ----------------
  umma::fence_before_sync();
  __syncthreads();
  if (tid < PINN_MAX_PDE) {
    double ssum = 0.0;
    for (int pq = 0; pq < T; ++pq) ssum += lred[tid * T + pq];
    p.lpart[(size_t)blockIdx.x * PINN_MAX_PDE + tid] = ssum;
  }
  if (warp == 0) umma::tmem_dealloc(tbase, CFG::tcTotal);
}

}  // namespace pinn
