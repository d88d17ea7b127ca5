// hallo_b200_attention, second-generation kernel: two 128-query tiles per CTA that ping-pong on the
// tensor core, P kept in TMEM (aliasing S), single-pass softmax with the whole S row in registers.
//
//   warp 0      TMA producer (Q_A, Q_B once; K_j / V_j through a 2-deep ring, shared by both tiles)
//   warp 1      MMA issuer:  S_t = Q_t K_j^T (SS)  and  O_t += P_t V_j (TS: A operand = P in TMEM)
//   warp 2      TMEM allocation;  warp 3 idle  (warpgroup 0 gives its registers away: setmaxnreg.dec 40)
//   warps 4-7   softmax of tile A   |  warps 8-11  softmax of tile B   (one thread per query row, 232 regs)
//
// Issue order on the tensor pipe is  ... PV_A(j) QK_A(j+1) | PV_B(j) QK_B(j+1) ...  so while one tile's
// softmax runs on the SFU / FMA pipes the other tile's MMAs run; tcgen05 ops of one thread execute in order,
// which is what makes overwriting S_t (and the P_t aliased onto it) by QK_t(j+1) safe after PV_t(j).
// Same operand layouts, segment (reference-KV concat) handling and lazy rescale as attn_tc.cu.
//
// Variants measured on B200 at L0 (C=320, L=4096, 32 frames, reference KV; profiles/r1_kbench_attn*.log):
//   this kernel (one MMA warp, BN=128, S single-buffered)                   2.40 ms  (XU pipe 61 %, ncu)
//   S double-buffered, BN=64                                                2.90 ms
//   + one MMA warp per tile                                                 2.62 ms  (BN=128: 3.20 ms)
//   + named-barrier ping-pong of the exp phases                             2.55 ms  (BN=128: 3.05 ms)
//   + 64/128 MUFU.EX2 issued back to back                                   2.96 ms  (BN=128: 3.60 ms)
// i.e. at head_dim 40 the step is bound by the SFU (16 ex2/clk/SM) plus per-step fixed latencies, and the
// simplest schedule won; the next lever is moving a share of the exponentials to the FMA pipe.
#include <type_traits>

#include "host_common.cuh"
#include "ptx.cuh"

namespace hb {

constexpr int kAttn2Threads = 384;   // warpgroup 0: TMA / MMA / TMEM-alloc / idle; warpgroups 1, 2: softmax of tile A, B
// MINB = 2 (two CTAs = four query tiles per SM, 64-key steps): no idle warps and no setmaxnreg -- ptxas allocates for
// the launch bound (65536 / 2 / 320 -> 96 registers per thread), which the 64-column S row fits
constexpr int kAttn2ThreadsOcc2 = 320;   // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 / 6-9 softmax of tile A / B

template <int D, int BN, int NSTG = 2>
struct Attn2Cfg {
  static constexpr int kChunks = (D + 63) / 64;
  static constexpr int kKSteps = (D + 15) / 16;
  static constexpr int kDv = ((D + 15) / 16) * 16;
  static constexpr int kQBytes = kChunks * 128 * 128;     // one Q tile
  static constexpr int kKVBytes = kChunks * BN * 128;
  static constexpr int kStages = NSTG;          // K / V ring depth (shared by both tiles of the CTA)
  static constexpr int kOffK = 2 * kQBytes;
  static constexpr int kOffV = kOffK + kStages * kKVBytes;
  static constexpr int kOffBar = kOffV + kStages * kKVBytes;
  static constexpr int kTotal = kOffBar + 256 + 1024;
  static constexpr uint32_t kSCol0 = 0, kSCol1 = BN;          // S_A, S_B (P_t aliases the first BN/2 columns)
  static constexpr uint32_t kOCol0 = 2 * BN;
  static constexpr uint32_t kOCol1 = 2 * BN + ((kDv + 31) / 32) * 32;
  static constexpr uint32_t kTmemCols = (kOCol1 + kDv) <= 256 ? 256 : 512;
  static_assert(kOCol1 + kDv <= 512, "TMEM budget");
};

// POLY: every POLY-th exponential of a row goes to the FMA pipe (exp2_poly) instead of the SFU; 0 = all on the SFU
template <typename T, int D, int BN, int POLY, int MINB = 1, int NSTG = 2>
__global__ void __launch_bounds__(MINB == 1 ? kAttn2Threads : kAttn2ThreadsOcc2, MINB)
attn2_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
                const __grid_constant__ CUtensorMap tmV1, const AttnDev p) {
  pdl_wait();
  using CF = Attn2Cfg<D, BN, NSTG>;
  constexpr int STAGES = CF::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CF::kOffBar);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // STAGES
  uint64_t* k_empty = k_full + STAGES;
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* s_full = v_empty + STAGES;     // 2 (tile A, tile B)
  uint64_t* p_full = s_full + 2;           // 2
  uint64_t* o_done = p_full + 2;           // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x;               // pair of query tiles: rows [qt*256, qt*256 + 256)
  const int head = blockIdx.y;
  const int frame = p.frames - 1 - (int)blockIdx.z;
  const int ref = (p.ref_index != nullptr) ? p.ref_index[frame] : -1;
  const int tiles_per_seg = (p.L + BN - 1) / BN;
  const int ntiles = tiles_per_seg * (ref >= 0 ? 2 : 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK0);
    tma_prefetch_desc(&tmV0);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&o_done[t], 1);
    }
    fence_barrier_init();
  }
  constexpr int W0 = (MINB == 1) ? 4 : 2;          // first softmax warp
  constexpr int kAllocWarp = (MINB == 1) ? 2 : 1;
  if (warp == kAllocWarp) tmem_alloc<CF::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  pdl_launch();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < W0) {
    if (MINB == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * CF::kQBytes);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int c = 0; c < CF::kChunks; ++c)
          tma_load_4d(smem + t * CF::kQBytes + c * (128 * 128), &tmQ, q_full, c * 64, head,
                      qt * 256 + t * 128, frame);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < ntiles; ++j) {
        const int seg = j / tiles_per_seg;
        const int kt = j - seg * tiles_per_seg;
        const CUtensorMap* mk = seg == 0 ? &tmK0 : &tmK1;
        const CUtensorMap* mv = seg == 0 ? &tmV0 : &tmV1;
        const int fr = seg == 0 ? frame : ref;
        mbar_wait(&k_empty[stage], phase ^ 1, 0x71);
        mbar_arrive_expect_tx(&k_full[stage], CF::kKVBytes);
#pragma unroll
        for (int c = 0; c < CF::kChunks; ++c)
          tma_load_4d(smem + CF::kOffK + stage * CF::kKVBytes + c * (BN * 128), mk, &k_full[stage],
                      c * 64, head, kt * BN, fr);
        mbar_wait(&v_empty[stage], phase ^ 1, 0x72);
        mbar_arrive_expect_tx(&v_full[stage], CF::kKVBytes);
#pragma unroll
        for (int c = 0; c < CF::kChunks; ++c)
          tma_load_4d(smem + CF::kOffV + stage * CF::kKVBytes + c * (BN * 128), mv, &v_full[stage],
                      c * 64, head, kt * BN, fr);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_qk = make_idesc_f16(128, BN, Cvt<T>::kFmt, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_f16(128, CF::kDv, Cvt<T>::kFmt, 0, 1);
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK = smem_u32(smem + CF::kOffK);
    const uint32_t sV = smem_u32(smem + CF::kOffV);
    const uint32_t s_col[2] = {tmem_base + CF::kSCol0, tmem_base + CF::kSCol1};
    const uint32_t o_col[2] = {tmem_base + CF::kOCol0, tmem_base + CF::kOCol1};

    auto issue_qk = [&](int t, int stage) {
      const uint32_t qbase = sQ + t * CF::kQBytes;
      const uint32_t kbase = sK + stage * CF::kKVBytes;
#pragma unroll
      for (int k = 0; k < CF::kKSteps; ++k) {
        const uint32_t off_q = (k >> 2) * (128 * 128) + (k & 3) * 32;
        const uint32_t off_k = (k >> 2) * (BN * 128) + (k & 3) * 32;
        umma_f16_ss(s_col[t], make_desc_sw128(qbase + off_q, 16, 1024),
                    make_desc_sw128(kbase + off_k, 16, 1024), idesc_qk, k != 0);
      }
    };

    mbar_wait(q_full, 0, 0x81);
    mbar_wait(&k_full[0], 0, 0x82);
    tc_fence_after();
    if (lane == 0) {
      issue_qk(0, 0);
      umma_commit(&s_full[0]);
      issue_qk(1, 0);
      umma_commit(&s_full[1]);
      umma_commit(&k_empty[0]);
    }
    __syncwarp();
    int kstage = 1 % STAGES, vstage = 0;
    uint32_t kphase = (STAGES == 1) ? 1 : 0, vphase = 0;

    for (int j = 0; j < ntiles; ++j) {
      const bool more = (j + 1 < ntiles);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_full[t], j & 1, 0x83);
        if (t == 0) mbar_wait(&v_full[vstage], vphase, 0x84);
        if (more && t == 0) mbar_wait(&k_full[kstage], kphase, 0x85);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t vbase = sV + vstage * CF::kKVBytes;
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) {
            // A = P_t in TMEM: 16 keys = 8 packed 32-bit columns per K step
            umma_f16_ts(o_col[t], s_col[t] + k * 8, make_desc_sw128(vbase + k * 2048, BN * 128, 1024),
                        idesc_pv, (j | k) != 0);
          }
          if (t == 1) umma_commit(&v_empty[vstage]);
          if (!more) umma_commit(&o_done[t]);
          if (more) {
            issue_qk(t, kstage);
            umma_commit(&s_full[t]);
            if (t == 1) umma_commit(&k_empty[kstage]);
          }
        }
        __syncwarp();
      }
      if (++vstage == STAGES) { vstage = 0; vphase ^= 1; }
      if (more) {
        if (++kstage == STAGES) { kstage = 0; kphase ^= 1; }
      }
    }
  }
  } else {
    if (MINB == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    // ============================ softmax warps ============================
    const int t = (warp - W0) >> 2;                      // query tile of this group of four warps
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + (t == 0 ? CF::kSCol0 : CF::kSCol1);
    const uint32_t o_addr = tmem_base + lane_addr + (t == 0 ? CF::kOCol0 : CF::kOCol1);
    float m_ref = -INFINITY;
    float l_sum = 0.f;

    for (int j = 0; j < ntiles; ++j) {
      const int kt = j % tiles_per_seg;
      const int key0 = kt * BN;
      mbar_wait(&s_full[t], j & 1, 0x91);
      tc_fence_after();
      uint32_t s[BN];
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) tmem_ld_x32(s_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_ld_wait();
      const bool tail = (key0 + BN > p.L);
      float mx = -INFINITY;
      if (!tail) {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains, 3-input max
#pragma unroll
        for (int i = 0; i < BN; i += 8) {
#pragma unroll
          for (int q = 0; q < 4; ++q) m4[q] = max3(m4[q], __uint_as_float(s[i + 2 * q]), __uint_as_float(s[i + 2 * q + 1]));
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
#pragma unroll
        for (int i = 0; i < BN; ++i)
          if (key0 + i < p.L) mx = fmaxf(mx, __uint_as_float(s[i]));
      }
      mx *= p.scale_log2;
      const bool need = (mx > m_ref + 8.0f);
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_ref, mx);
        if (j > 0) {
          // s_full(j) was committed after P_t V_{j-1}: O_t is quiescent here (in-order tensor pipe)
          const float f = fast_exp2(m_ref - m_new);
#pragma unroll
          for (int c = 0; c < CF::kDv / 8; ++c) {
            uint32_t r[8];
            tmem_ld_x8(o_addr + c * 8, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
            tmem_st_x8(o_addr + c * 8, r);
          }
          l_sum *= f;
        }
        m_ref = m_new;
      }
      float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < BN; i += 2) {
        const float x0 = fmaf(__uint_as_float(s[i]), p.scale_log2, -m_ref);
        const float x1 = fmaf(__uint_as_float(s[i + 1]), p.scale_log2, -m_ref);
        constexpr int PM = POLY > 0 ? POLY : 1;
        float e0 = (POLY > 0 && (i % PM) == PM - 1) ? exp2_poly(x0) : fast_exp2(x0);
        float e1 = (POLY > 0 && ((i + 1) % PM) == PM - 1) ? exp2_poly(x1) : fast_exp2(x1);
        if (tail) {
          if (key0 + i >= p.L) e0 = 0.f;
          if (key0 + i + 1 >= p.L) e1 = 0.f;
        }
        ps4[(i >> 1) & 3] += e0 + e1;
        s[i >> 1] = Cvt<T>::pack2(e0, e1);       // P packed two keys per 32-bit TMEM column
      }
      l_sum += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
#pragma unroll
      for (int c = 0; c < BN / 64; ++c) tmem_st_x32(s_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }


    // ---- epilogue: O / l -> global ----
    mbar_wait(&o_done[t], 0, 0x92);
    tc_fence_after();
    const float inv = 1.0f / l_sum;
    const int qrow = qt * 256 + t * 128 + row;
    T* out = reinterpret_cast<T*>(p.O) + ((long long)frame * p.L + qrow) * p.ldo + head * D;
#pragma unroll
    for (int c = 0; c < CF::kDv / 8; ++c) {
      uint32_t r[8];
      tmem_ld_x8(o_addr + c * 8, r);
      tmem_ld_wait();
      if (c * 8 < D && qrow < p.L) {
        uint4 o4;
        o4.x = Cvt<T>::pack2(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
        o4.y = Cvt<T>::pack2(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
        o4.z = Cvt<T>::pack2(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
        o4.w = Cvt<T>::pack2(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
        *reinterpret_cast<uint4*>(out + c * 8) = o4;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc<CF::kTmemCols>(tmem_base);
  }
}

template <typename T, int D, int BN, int POLY, int MINB = 1, int NSTG = 2>
static int launch_attn2(const hb_attention_params* q, cudaStream_t stream) {
  using CF = Attn2Cfg<D, BN, NSTG>;
  static_assert(CF::kTotal <= 232448, "attention v2 smem budget");
  CUtensorMap tmQ, tmK0, tmV0, tmK1, tmV1;
  int rc;
  if ((rc = make_qkv_map(&tmQ, q->dtype, q->Q, D, q->heads, q->L, q->frames, q->ldq, 128))) return rc;
  if ((rc = make_qkv_map(&tmK0, q->dtype, q->K, D, q->heads, q->L, q->frames, q->ldk, BN))) return rc;
  if ((rc = make_qkv_map(&tmV0, q->dtype, q->V, D, q->heads, q->L, q->frames, q->ldv, BN))) return rc;
  if (q->ref_index != nullptr) {
    if (q->Kref == nullptr || q->Vref == nullptr || q->ref_frames <= 0)
      return fail(HB_ERR_NULL, "attention: ref_index given without Kref/Vref");
    if ((rc = make_qkv_map(&tmK1, q->dtype, q->Kref, D, q->heads, q->L, q->ref_frames, q->ldkref, BN))) return rc;
    if ((rc = make_qkv_map(&tmV1, q->dtype, q->Vref, D, q->heads, q->L, q->ref_frames, q->ldvref, BN))) return rc;
  } else {
    tmK1 = tmK0;
    tmV1 = tmV0;
  }
  AttnDev d{};
  d.L = q->L;
  d.heads = q->heads;
  d.frames = q->frames;
  d.ref_index = q->ref_index;
  d.O = q->O;
  d.ldo = q->ldo;
  d.scale_log2 = (float)(1.4426950408889634 / sqrt((double)D));
  auto kern = attn2_tc_kernel<T, D, BN, POLY, MINB, NSTG>;
  static bool attr_set = false;
  if (!attr_set) {
    HB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::kTotal));
    attr_set = true;
  }
  dim3 grid((q->L + 255) / 256, q->heads, q->frames);
  launch_kernel(kern, grid, MINB == 1 ? kAttn2Threads : kAttn2ThreadsOcc2, CF::kTotal, stream, tmQ, tmK0, tmV0, tmK1, tmV1, d);
  HB_LAUNCH_CHECK();
  return HB_OK;
}

}  // namespace hb
