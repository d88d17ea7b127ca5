// hallo_b200_attention, head_dim 40, "half-split" schedule (option attn_split, opt-in until its A/B run is green):
// the 64-key score tile of attn2_tc_kernel<.., 64, .., MINB = 2> is handed between the softmax warps and the tensor
// core in two 32-key halves, each with its own S-ready / P-ready barriers.
//
// Why: in attn2 the softmax warps of a tile finish P(j), then wait for the chain  p_full -> PV(j) -> QK(j+1) -> s_full
// -> tcgen05.ld  before they can start step j+1 -- 31 % of their samples sit on s_full (profiles/r2_ncu_attn_stalls.txt)
// and the SFU, the unit that bounds the kernel, idles 26 % of the time.  Here half h of step j+1 is computed
// (QK_h(j+1), issued right after PV_h(j)) while the softmax warps are still busy with the OTHER half of step j, so S is
// (almost) always ready: double buffering without a second S buffer -- TMEM stays at 224 columns, two CTAs per SM.
//
//   per tile t, TMEM columns:  S_h at s_col + 32 h (fp32, 32 keys);  P_h (fp16 pairs, 16 columns) aliases the start of S_h
//   MMA warp, per K/V step j:  for h in {0,1}: for t in {0,1}:  wait p_full[t][h](j);  O_t += P_h V[32h..32h+31];
//                                                             S_h(j+1) = Q_t K(j+1)[32h..]^T;  commit s_full[t][h]
//   softmax warps of tile t:   for h in {0,1}:  wait s_full[t][h](j);  tcgen05.ld 32 columns;  online max (lazy rescale
//                              of O behind pv_bar[t]);  32 x ex2;  P_h -> TMEM;  arrive p_full[t][h]
// Everything else (TMA ring, operand layouts, reference-KV segment, epilogue) is attn2's.
#include "host_common.cuh"
#include "ptx.cuh"

namespace hb {

constexpr int kAttn4Threads = 320;   // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 / 6-9 softmax of tile A / B
constexpr int kA4BN = 64, kA4H = 32, kA4D = 40;

struct Attn4Cfg {
  static constexpr int kKSteps = 3;                      // 40 -> 48 (TMA zero fill), K = 16 per MMA
  static constexpr int kDv = 48;
  static constexpr int kQBytes = 128 * 128;              // one Q tile (one 64-column swizzle chunk)
  static constexpr int kKVBytes = kA4BN * 128;
  static constexpr int kStages = 2;
  static constexpr int kOffK = 2 * kQBytes;
  static constexpr int kOffV = kOffK + kStages * kKVBytes;
  static constexpr int kOffBar = kOffV + kStages * kKVBytes;
  static constexpr int kTotal = kOffBar + 256 + 1024;
  static constexpr uint32_t kSCol0 = 0, kSCol1 = kA4BN;
  static constexpr uint32_t kOCol0 = 2 * kA4BN, kOCol1 = 2 * kA4BN + 64;
  static constexpr uint32_t kTmemCols = 256;
};

template <typename T>
__global__ void __launch_bounds__(kAttn4Threads, 2)
attn4_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
                const __grid_constant__ CUtensorMap tmV1, const AttnDev p) {
  using CF = Attn4Cfg;
  constexpr int STAGES = CF::kStages, BN = kA4BN, HB = kA4H;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CF::kOffBar);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // STAGES
  uint64_t* k_empty = k_full + STAGES;
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* s_full = v_empty + STAGES;     // [tile][half] = 4
  uint64_t* p_full = s_full + 4;           // 4
  uint64_t* pv_bar = p_full + 4;           // [tile][half] = 4: completes when P_h V of step j is done -> phase j.  One
                                           // barrier per half keeps a waiter within one phase of it (a parity wait cannot
                                           // tell phase j from phase j + 2); guards the lazy rescale of O
  uint64_t* o_done = pv_bar + 4;           // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x;
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
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_bar[i], 1);
    }
    for (int t = 0; t < 2; ++t) mbar_init(&o_done[t], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<CF::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer (as attn2) ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * CF::kQBytes);
#pragma unroll
      for (int t = 0; t < 2; ++t)
        tma_load_4d(smem + t * CF::kQBytes, &tmQ, q_full, 0, head, qt * 256 + t * 128, frame);
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
        tma_load_4d(smem + CF::kOffK + stage * CF::kKVBytes, mk, &k_full[stage], 0, head, kt * BN, fr);
        mbar_wait(&v_empty[stage], phase ^ 1, 0x72);
        mbar_arrive_expect_tx(&v_full[stage], CF::kKVBytes);
        tma_load_4d(smem + CF::kOffV + stage * CF::kKVBytes, mv, &v_full[stage], 0, head, kt * BN, fr);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_qk = make_idesc_f16(128, HB, Cvt<T>::kFmt, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_f16(128, CF::kDv, Cvt<T>::kFmt, 0, 1);
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK = smem_u32(smem + CF::kOffK);
    const uint32_t sV = smem_u32(smem + CF::kOffV);
    const uint32_t s_col[2] = {tmem_base + CF::kSCol0, tmem_base + CF::kSCol1};
    const uint32_t o_col[2] = {tmem_base + CF::kOCol0, tmem_base + CF::kOCol1};

    // S_h = Q_t K[32h .. 32h+31]^T : key row r of the K tile sits at r * 128 B (8-row swizzle atoms of 1024 B)
    auto issue_qk = [&](int t, int h, int stage) {
      const uint32_t qbase = sQ + t * CF::kQBytes;
      const uint32_t kbase = sK + stage * CF::kKVBytes + h * (HB * 128);
#pragma unroll
      for (int k = 0; k < CF::kKSteps; ++k)
        umma_f16_ss(s_col[t] + h * HB, make_desc_sw128(qbase + k * 32, 16, 1024), make_desc_sw128(kbase + k * 32, 16, 1024),
                    idesc_qk, k != 0);
    };

    mbar_wait(q_full, 0, 0x81);
    mbar_wait(&k_full[0], 0, 0x82);
    tc_fence_after();
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          issue_qk(t, h, 0);
          umma_commit(&s_full[t * 2 + h]);
        }
      umma_commit(&k_empty[0]);
    }
    __syncwarp();
    int kstage = 1 % STAGES, vstage = 0;
    uint32_t kphase = (STAGES == 1) ? 1 : 0, vphase = 0;
    for (int j = 0; j < ntiles; ++j) {
      const bool more = (j + 1 < ntiles);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t * 2 + h], j & 1, 0x83);
          if (h == 0 && t == 0) {
            mbar_wait(&v_full[vstage], vphase, 0x84);
            if (more) mbar_wait(&k_full[kstage], kphase, 0x85);
          }
          tc_fence_after();
          if (lane == 0) {
            const uint32_t vbase = sV + vstage * CF::kKVBytes;
#pragma unroll
            for (int k = 0; k < HB / 16; ++k)     // A = P_h in TMEM (8 packed columns per 16 keys), B = V rows 32h + 16k ..
              umma_f16_ts(o_col[t], s_col[t] + h * HB + k * 8, make_desc_sw128(vbase + (2 * h + k) * 2048, BN * 128, 1024),
                          idesc_pv, (j | h | k) != 0);
            umma_commit(&pv_bar[t * 2 + h]);
            if (h == 1 && t == 1) umma_commit(&v_empty[vstage]);
            if (!more && h == 1) umma_commit(&o_done[t]);
            if (more) {
              issue_qk(t, h, kstage);
              umma_commit(&s_full[t * 2 + h]);
              if (h == 1 && t == 1) umma_commit(&k_empty[kstage]);
            }
          }
          __syncwarp();
        }
      }
      if (++vstage == STAGES) { vstage = 0; vphase ^= 1; }
      if (more) {
        if (++kstage == STAGES) { kstage = 0; kphase ^= 1; }
      }
    }
  } else {
    // ============================ softmax warps ============================
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + (t == 0 ? CF::kSCol0 : CF::kSCol1);
    const uint32_t o_addr = tmem_base + lane_addr + (t == 0 ? CF::kOCol0 : CF::kOCol1);
    float m_ref = -INFINITY;
    float l_sum = 0.f;

    for (int j = 0; j < ntiles; ++j) {
      const int kt = j % tiles_per_seg;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int key0 = kt * BN + h * HB;
        // consume the completion of P_h V(j-1) (long done: it was requested two halves ago) so that this warp never
        // falls more than one phase behind pv_bar[t][h]
        if (j >= 1) mbar_wait(&pv_bar[t * 2 + h], (uint32_t)((j - 1) & 1), 0x94);
        mbar_wait(&s_full[t * 2 + h], j & 1, 0x91);
        tc_fence_after();
        uint32_t s[HB];
        tmem_ld_x32(s_addr + h * HB, s);
        tmem_ld_wait();
        const bool tail = (key0 + HB > p.L);
        float mx = -INFINITY;
        if (!tail) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < HB; i += 8) {
#pragma unroll
            for (int q = 0; q < 4; ++q) m4[q] = max3(m4[q], __uint_as_float(s[i + 2 * q]), __uint_as_float(s[i + 2 * q + 1]));
          }
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        } else {
#pragma unroll
          for (int i = 0; i < HB; ++i)
            if (key0 + i < p.L) mx = fmaxf(mx, __uint_as_float(s[i]));
        }
        mx *= p.scale_log2;
        // keys past L in a ragged last tile: the whole half may be out of range (mx = -inf): nothing to do for the max
        const bool need = (mx > m_ref + 8.0f);
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = fmaxf(m_ref, mx);
          if (j > 0 || h > 0) {
            // O_t is touched by every P V group already requested: wait for the most recent one -- half 0 of this
            // step, or half 1 of the previous step
            if (h == 1) mbar_wait(&pv_bar[t * 2 + 0], (uint32_t)(j & 1), 0x95);
            else mbar_wait(&pv_bar[t * 2 + 1], (uint32_t)((j - 1) & 1), 0x95);
            tc_fence_after();
            const float f = (m_ref == -INFINITY) ? 0.f : fast_exp2(m_ref - m_new);
#pragma unroll
            for (int cc = 0; cc < CF::kDv / 8; ++cc) {
              uint32_t r[8];
              tmem_ld_x8(o_addr + cc * 8, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
              tmem_st_x8(o_addr + cc * 8, r);
            }
            l_sum *= f;
          }
          m_ref = m_new;
        }
        float ps4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[HB / 2];
#pragma unroll
        for (int i = 0; i < HB; i += 2) {
          float e0 = fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, -m_ref));
          float e1 = fast_exp2(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, -m_ref));
          if (tail) {
            if (key0 + i >= p.L) e0 = 0.f;
            if (key0 + i + 1 >= p.L) e1 = 0.f;
          }
          ps4[(i >> 1) & 3] += e0 + e1;
          pk[i >> 1] = Cvt<T>::pack2(e0, e1);
        }
        l_sum += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
        tmem_st_x16(s_addr + h * HB, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t * 2 + h]);
      }
    }

    // ---- epilogue: O / l -> global ----
    mbar_wait(&o_done[t], 0, 0x92);
    tc_fence_after();
    const float inv = 1.0f / l_sum;
    const int qrow = qt * 256 + t * 128 + row;
    T* out = reinterpret_cast<T*>(p.O) + ((long long)frame * p.L + qrow) * p.ldo + head * kA4D;
#pragma unroll
    for (int cc = 0; cc < CF::kDv / 8; ++cc) {
      uint32_t r[8];
      tmem_ld_x8(o_addr + cc * 8, r);
      tmem_ld_wait();
      if (cc * 8 < kA4D && qrow < p.L) {
        uint4 o4;
        o4.x = Cvt<T>::pack2(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
        o4.y = Cvt<T>::pack2(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
        o4.z = Cvt<T>::pack2(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
        o4.w = Cvt<T>::pack2(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
        *reinterpret_cast<uint4*>(out + cc * 8) = o4;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CF::kTmemCols>(tmem_base);
  }
}

template <typename T>
static int launch_attn4(const hb_attention_params* q, cudaStream_t stream) {
  using CF = Attn4Cfg;
  CUtensorMap tmQ, tmK0, tmV0, tmK1, tmV1;
  int rc;
  if ((rc = make_qkv_map(&tmQ, q->dtype, q->Q, kA4D, q->heads, q->L, q->frames, q->ldq, 128))) return rc;
  if ((rc = make_qkv_map(&tmK0, q->dtype, q->K, kA4D, q->heads, q->L, q->frames, q->ldk, kA4BN))) return rc;
  if ((rc = make_qkv_map(&tmV0, q->dtype, q->V, kA4D, q->heads, q->L, q->frames, q->ldv, kA4BN))) return rc;
  if (q->ref_index != nullptr) {
    if (q->Kref == nullptr || q->Vref == nullptr || q->ref_frames <= 0)
      return fail(HB_ERR_NULL, "attention: ref_index given without Kref/Vref");
    if ((rc = make_qkv_map(&tmK1, q->dtype, q->Kref, kA4D, q->heads, q->L, q->ref_frames, q->ldkref, kA4BN))) return rc;
    if ((rc = make_qkv_map(&tmV1, q->dtype, q->Vref, kA4D, q->heads, q->L, q->ref_frames, q->ldvref, kA4BN))) return rc;
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
  d.scale_log2 = (float)(1.4426950408889634 / sqrt((double)kA4D));
  auto kern = attn4_tc_kernel<T>;
  static bool attr_set = false;
  if (!attr_set) {
    HB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::kTotal));
    attr_set = true;
  }
  dim3 grid((q->L + 255) / 256, q->heads, q->frames);
  kern<<<grid, kAttn4Threads, CF::kTotal, stream>>>(tmQ, tmK0, tmV0, tmK1, tmV1, d);
  HB_LAUNCH_CHECK();
  return HB_OK;
}

}  // namespace hb
