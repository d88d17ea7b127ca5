// hallo_b200_attention, third generation (head_dim 40 only): S never touches TMEM.
//
// attn2_tc.cu is bound by two ~1030-clk phases per 128x128 score tile: the tcgen05.ld of S (64 KB at 64 B/clk) and the
// 16 K exponentials on the SFU.  Here Q K^T runs on the warp-level tensor-core path (mma.sync m16n8k16) with the
// accumulators in registers, so the TMEM read disappears; only P V stays on tcgen05:
//
//   warp 0      TMA producer (Q_A, Q_B once; K_j / V_j through a 2-deep ring, shared by both query tiles)
//   warp 1      tcgen05 issuer:  O_t += P_t V_j  (A = P_t in TMEM, 8 k-steps of 16 keys); writes the ones column
//   warp 2      TMEM allocation;  warp 3 idle    (warpgroup 0: setmaxnreg.dec 40)
//   warps 4-7   tile A, warps 8-11 tile B (setmaxnreg.inc 232): warp w owns rows 32w..32w+31 of its tile
//               S = Q K^T: Q fragments resident in registers, K fragments by ldmatrix from the 128B-swizzled TMA tile
//               online softmax on the C fragments (row max / lazy rescale per fragment row, quad shuffles)
//               P packed to 16-bit pairs IS the register layout of tcgen05.st.16x128b (lane (g,t): rows g, g+8,
//               32-bit column t), stored to the TMEM lanes of the warp's rows, double-buffered per tile
//
// The softmax denominator is accumulated by the tensor core: column 40 of every V row (TMA zero fill) is set to 1.0,
// so O[:, 40] = sum_j P (with the same rounded P that builds O) and no FADD is spent on row sums.
// Lazy rescale touches O in TMEM (32x32b load/store, row r in lane r) after waiting for the previous P V.
// Opt-in (option "attn_v3") -- never run on hardware; the fragment / TMEM-store index algebra is modelled in
// tests/test_kernel_logic_cpu.py, the barrier protocol in tests/test_protocol_sim_cpu.py.
#include <type_traits>

#include "host_common.cuh"
#include "ptx.cuh"

namespace hb {

constexpr int kAttn3Threads = 384;

struct Attn3Cfg {
  static constexpr int D = 40, BN = 128, kDv = 48, kStages = 2;
  static constexpr int kQBytes = 128 * 128;                 // one 128-row tile, 64 columns (40 real) x 2 B
  static constexpr int kKVBytes = BN * 128;
  static constexpr int kOffK = 2 * kQBytes;
  static constexpr int kOffV = kOffK + kStages * kKVBytes;
  static constexpr int kOffBar = kOffV + kStages * kKVBytes;
  static constexpr int kTotal = kOffBar + 256 + 1024;
  // TMEM columns: P[tile][buffer] 64 packed columns each, then O[tile] (48 used of 64)
  static constexpr uint32_t kPCol = 0, kPStride = 64, kOCol = 256, kOStride = 64, kTmemCols = 512;
};

// 16 TMEM lanes x (4 x N) 32-bit columns from the m16n8-fragment register layout: regs {row g, row g+8} per 8-key tile
__device__ __forceinline__ void tmem_st_16x128b_x8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <typename T, int POLY>
__global__ void __launch_bounds__(kAttn3Threads, 1)
attn3_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
                const __grid_constant__ CUtensorMap tmV1, const AttnDev p) {
  using CF = Attn3Cfg;
  constexpr int BN = CF::BN, D = CF::D, STAGES = CF::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CF::kOffBar);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // STAGES
  uint64_t* k_empty = k_full + STAGES;     // count 8: every compute warp has read its K fragments
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* p_full = v_empty + STAGES;     // [tile][buffer], count 4
  uint64_t* p_free = p_full + 4;           // [tile][buffer], tcgen05 commit after P V
  uint64_t* o_done = p_free + 4;           // [tile]
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
      mbar_init(&k_empty[s], 8);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&p_full[i], 4);
      mbar_init(&p_free[i], 1);
    }
    mbar_init(&o_done[0], 1);
    mbar_init(&o_done[1], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CF::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (warp == 0) {
      // ============================ TMA producer ============================
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * CF::kQBytes);
#pragma unroll
        for (int t = 0; t < 2; ++t) tma_load_4d(smem + t * CF::kQBytes, &tmQ, q_full, 0, head, qt * 256 + t * 128, frame);
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
      // ============================ tcgen05 issuer: O_t += P_t V_j ============================
      constexpr uint32_t idesc_pv = make_idesc_f16(128, CF::kDv, Cvt<T>::kFmt, 0, 1);
      const uint32_t sV = smem_u32(smem + CF::kOffV);
      int vstage = 0;
      uint32_t vphase = 0;
      for (int j = 0; j < ntiles; ++j) {
        const int buf = j & 1;
        const uint32_t bph = (j >> 1) & 1;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t == 0) {
            mbar_wait(&v_full[vstage], vphase, 0x84);
            // ones column: column 40 = 16-byte chunk 5 of the 128-byte row, 128B swizzle: chunk ^= row & 7
            const uint16_t one = std::is_same<T, __half>::value ? (uint16_t)0x3C00 : (uint16_t)0x3F80;
#pragma unroll
            for (int i = 0; i < BN / 32; ++i) {
              const uint32_t r = (uint32_t)lane + 32u * i;
              asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(sV + vstage * CF::kKVBytes + r * 128u + ((5u ^ (r & 7u)) << 4)),
                           "h"(one)
                           : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
          }
          mbar_wait(&p_full[t * 2 + buf], bph, 0x83);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t vbase = sV + vstage * CF::kKVBytes;
            const uint32_t p_col = tmem_base + CF::kPCol + (t * 2 + buf) * CF::kPStride;
            const uint32_t o_col = tmem_base + CF::kOCol + t * CF::kOStride;
#pragma unroll
            for (int k = 0; k < BN / 16; ++k)
              umma_f16_ts(o_col, p_col + k * 8, make_desc_sw128(vbase + k * 2048, BN * 128, 1024), idesc_pv, (j | k) != 0);
            umma_commit(&p_free[t * 2 + buf]);
            if (j + 1 == ntiles) umma_commit(&o_done[t]);
            if (t == 1) umma_commit(&v_empty[vstage]);
          }
          __syncwarp();
        }
        if (++vstage == STAGES) {
          vstage = 0;
          vphase ^= 1;
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    // ============================ compute warps ============================
    const int t = (warp - 4) >> 2;                 // query tile of this warpgroup
    const int wq = warp & 3;                       // TMEM lane quarter = rows 32 wq .. 32 wq + 31 of the tile
    const int tq = lane & 3;                      // C fragment: rows (lane >> 2), +8; columns 2 tq, 2 tq + 1
    const uint32_t sQ = smem_u32(smem + t * CF::kQBytes);
    const uint32_t sK = smem_u32(smem + CF::kOffK);
    const uint32_t lane_addr = ((uint32_t)(wq * 32)) << 16;
    const uint32_t o_addr = tmem_base + lane_addr + CF::kOCol + t * CF::kOStride;     // 32x32b view: row = lane

    // Q fragments (A operand), resident: [m tile][k step][4]
    uint32_t qa[2][3][4];
    mbar_wait(q_full, 0, 0x90);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        const uint32_t row = (uint32_t)(wq * 32 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8);
        const uint32_t chunk = (uint32_t)(2 * ks + (lane >> 4));
        ldsm_x4(sQ + row * 128u + ((chunk ^ (row & 7u)) << 4), qa[mt][ks]);
      }

    float m_ref[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}};   // [m tile][row g / g + 8], log2 units
    int kstage = 0;
    uint32_t kphase = 0;

    for (int j = 0; j < ntiles; ++j) {
      const int kt = j % tiles_per_seg;
      const int key0 = kt * BN;
      const bool tail = (key0 + BN > p.L);
      const int buf = j & 1;

      // ---- S = Q K^T on mma.sync, accumulators in registers ----
      float s[2][16][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 16; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
      mbar_wait(&k_full[kstage], kphase, 0x91);
      const uint32_t kbase = sK + kstage * CF::kKVBytes;
#pragma unroll
      for (int np = 0; np < 8; ++np) {              // pairs of 8-key tiles
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          // x4: (keys 0-7, dims 0-7) (keys 0-7, dims 8-15) (keys 8-15, dims 0-7) (keys 8-15, dims 8-15)
          uint32_t kb[4];
          const uint32_t key = (uint32_t)(np * 16 + ((lane >> 4) & 1) * 8 + (lane & 7));
          const uint32_t chunk = (uint32_t)(2 * ks + ((lane >> 3) & 1));
          ldsm_x4(kbase + key * 128u + ((chunk ^ (key & 7u)) << 4), kb);
          const uint32_t b0[2] = {kb[0], kb[1]}, b1[2] = {kb[2], kb[3]};
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            WarpMma<T>::mma(s[mt][2 * np], qa[mt][ks], b0);
            WarpMma<T>::mma(s[mt][2 * np + 1], qa[mt][ks], b1);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&k_empty[kstage]);
      if (++kstage == STAGES) {
        kstage = 0;
        kphase ^= 1;
      }

      // ---- row maxima (fragment rows), lazy rescale ----
      float mx[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float a = -INFINITY, b = -INFINITY;
          if (!tail) {
#pragma unroll
            for (int nt = 0; nt < 16; nt += 2) {
              a = max3(a, s[mt][nt][hh * 2], s[mt][nt][hh * 2 + 1]);
              b = max3(b, s[mt][nt + 1][hh * 2], s[mt][nt + 1][hh * 2 + 1]);
            }
          } else {
#pragma unroll
            for (int nt = 0; nt < 16; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e)
                if (key0 + nt * 8 + 2 * tq + e < p.L) a = fmaxf(a, s[mt][nt][hh * 2 + e]);
          }
          float v = fmaxf(a, b);
          v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
          v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
          mx[mt][hh] = v * p.scale_log2;
        }
      bool need = false;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) need = need || (mx[mt][hh] > m_ref[mt][hh] + 8.0f);
      if (__any_sync(0xffffffffu, need)) {
        float f[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const float m_new = fmaxf(m_ref[mt][hh], mx[mt][hh]);
            f[mt][hh] = (m_ref[mt][hh] == -INFINITY) ? 0.f : fast_exp2(m_ref[mt][hh] - m_new);   // first step: O is empty
            if (mx[mt][hh] > m_ref[mt][hh] + 8.0f) m_ref[mt][hh] = m_new; else f[mt][hh] = 1.0f;
          }
        if (j > 0) {
          // O_t must be quiescent: wait for P V of step j-1 (buffer (j-1)&1, its ((j-1)>>1)-th use)
          mbar_wait(&p_free[t * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1, 0x93);
          tc_fence_after();
          // the factor of row 32 wq + L lives in quad (L & 7) as f[L >> 4][(L >> 3) & 1]
          const int src = (lane & 7) * 4;
          const float f00 = __shfl_sync(0xffffffffu, f[0][0], src), f01 = __shfl_sync(0xffffffffu, f[0][1], src);
          const float f10 = __shfl_sync(0xffffffffu, f[1][0], src), f11 = __shfl_sync(0xffffffffu, f[1][1], src);
          const float fr = (lane & 16) ? ((lane & 8) ? f11 : f10) : ((lane & 8) ? f01 : f00);
#pragma unroll
          for (int c = 0; c < CF::kDv / 8; ++c) {
            uint32_t r[8];
            tmem_ld_x8(o_addr + c * 8, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * fr);
            tmem_st_x8(o_addr + c * 8, r);
          }
          tmem_st_wait();
          tc_fence_before();
        }
      }

      // ---- exponentials, P packed in place: pk[mt][nt*2 + hh] = (row g + 8 hh, keys 8 nt + 2 tq, +1) ----
      uint32_t pk[2][32];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 16; ++nt)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const float x0 = fmaf(s[mt][nt][hh * 2], p.scale_log2, -m_ref[mt][hh]);
            const float x1 = fmaf(s[mt][nt][hh * 2 + 1], p.scale_log2, -m_ref[mt][hh]);
            constexpr int PM = POLY > 0 ? POLY : 1;
            float e0 = (POLY > 0 && (nt % PM) == PM - 1) ? exp2_poly(x0) : fast_exp2(x0);
            float e1 = (POLY > 0 && (nt % PM) == PM - 1) ? exp2_poly(x1) : fast_exp2(x1);
            if (tail) {
              if (key0 + nt * 8 + 2 * tq >= p.L) e0 = 0.f;
              if (key0 + nt * 8 + 2 * tq + 1 >= p.L) e1 = 0.f;
            }
            pk[mt][nt * 2 + hh] = Cvt<T>::pack2(e0, e1);
          }

      // ---- P -> TMEM (buffer j & 1 of this tile), free once P V of step j-2 has completed ----
      mbar_wait(&p_free[t * 2 + buf], ((j >> 1) & 1) ^ 1, 0x94);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t p_addr = tmem_base + (((uint32_t)(wq * 32 + mt * 16)) << 16) + CF::kPCol + (t * 2 + buf) * CF::kPStride;
        tmem_st_16x128b_x8(p_addr, *reinterpret_cast<uint32_t(*)[16]>(&pk[mt][0]));          // keys 0..63 -> columns 0..31
        tmem_st_16x128b_x8(p_addr + 32, *reinterpret_cast<uint32_t(*)[16]>(&pk[mt][16]));    // keys 64..127
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t * 2 + buf]);
    }

    // ---- epilogue: O / O[:, 40] -> global ----
    mbar_wait(&o_done[t], 0, 0x92);
    tc_fence_after();
    float inv;
    {
      uint32_t r[8];
      tmem_ld_x8(o_addr + 40, r);
      tmem_ld_wait();
      inv = 1.0f / __uint_as_float(r[0]);
    }
    const int row = wq * 32 + lane;
    const int qrow = qt * 256 + t * 128 + row;
    T* out = reinterpret_cast<T*>(p.O) + ((long long)frame * p.L + qrow) * p.ldo + head * D;
#pragma unroll
    for (int c = 0; c < D / 8; ++c) {
      uint32_t r[8];
      tmem_ld_x8(o_addr + c * 8, r);
      tmem_ld_wait();
      if (qrow < p.L) {
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
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<CF::kTmemCols>(tmem_base);
  }
}

template <typename T, int POLY>
static int launch_attn3(const hb_attention_params* q, cudaStream_t stream) {
  using CF = Attn3Cfg;
  static_assert(CF::kTotal <= 232448, "attention v3 smem budget");
  CUtensorMap tmQ, tmK0, tmV0, tmK1, tmV1;
  int rc;
  if ((rc = make_qkv_map(&tmQ, q->dtype, q->Q, CF::D, q->heads, q->L, q->frames, q->ldq, 128))) return rc;
  if ((rc = make_qkv_map(&tmK0, q->dtype, q->K, CF::D, q->heads, q->L, q->frames, q->ldk, CF::BN))) return rc;
  if ((rc = make_qkv_map(&tmV0, q->dtype, q->V, CF::D, q->heads, q->L, q->frames, q->ldv, CF::BN))) return rc;
  if (q->ref_index != nullptr) {
    if (q->Kref == nullptr || q->Vref == nullptr || q->ref_frames <= 0)
      return fail(HB_ERR_NULL, "attention: ref_index given without Kref/Vref");
    if ((rc = make_qkv_map(&tmK1, q->dtype, q->Kref, CF::D, q->heads, q->L, q->ref_frames, q->ldkref, CF::BN))) return rc;
    if ((rc = make_qkv_map(&tmV1, q->dtype, q->Vref, CF::D, q->heads, q->L, q->ref_frames, q->ldvref, CF::BN))) return rc;
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
  d.scale_log2 = (float)(1.4426950408889634 / sqrt((double)CF::D));
  auto kern = attn3_tc_kernel<T, POLY>;
  static bool attr_set = false;
  if (!attr_set) {
    HB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::kTotal));
    attr_set = true;
  }
  dim3 grid((q->L + 255) / 256, q->heads, q->frames);
  kern<<<grid, kAttn3Threads, CF::kTotal, stream>>>(tmQ, tmK0, tmV0, tmK1, tmV1, d);
  HB_LAUNCH_CHECK();
  return HB_OK;
}

}  // namespace hb
