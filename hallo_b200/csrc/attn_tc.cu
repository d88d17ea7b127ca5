// hallo_b200_attention: fused softmax(Q K^T / sqrt(d)) V on tcgen05 tensor cores.
//
// One CTA = one (frame, head, 128-query tile).  Keys come in up to two SEGMENTS that are
// concatenated in-kernel: segment 0 = the frame's own tokens, segment 1 (optional, per frame) =
// the ReferenceNet tokens of CFG half ref_index[frame] -- this is the `torch.cat([norm_hidden_states,
// bank])` of mutual_self_attention.py:253-263, without ever materialising the 2L-key tensor, and
// without recomputing the uncond half (mutual_self_attention.py:264-284): uncond frames simply have
// ref_index = -1.
//
//   warp 0     TMA producer: Q tile once, then K / V tiles of BN keys through a STAGES-deep ring
//   warp 1     MMA issuer:   S[j&1] = Q K_j^T  (TMEM, fp32)   and   O += P_j V_j  (TMEM, fp32)
//   warps 2-5  softmax:      one thread per query row; S from TMEM, online max / sum in fp32,
//                            P_j (fp16/bf16) written 128B-swizzled to smem as the A operand of PV;
//                            O is only rescaled when the running max moved by > 2^8 (lazy rescale).
//
// Head dims 40 / 80 / 160 are not multiples of the 64-element swizzle span: TMA boxes of 64 columns
// over a (d, head, token, frame) tensor map zero-fill the columns >= d, so Q/K/V stay unpadded in HBM.
#include "host_common.cuh"
#include "ptx.cuh"

namespace hb {

constexpr int kAttnThreads = 192;

struct AttnDev {
  int L;              // queries per frame == keys per segment
  int heads;
  int frames;
  const int* ref_index;  // [frames] or nullptr
  void* O;
  long long ldo;
  float scale_log2;   // d^-0.5 * log2(e)
};

template <int D, int BN, int STAGES>
struct AttnCfg {
  static constexpr int kChunks = (D + 63) / 64;          // 64-column swizzle chunks per row
  static constexpr int kKSteps = (D + 15) / 16;          // MMA K steps for Q K^T
  static constexpr int kDv = ((D + 15) / 16) * 16;       // N of the P V MMA (48 / 80 / 160)
  static constexpr int kQBytes = kChunks * 128 * 128;
  static constexpr int kKVBytes = kChunks * BN * 128;    // one K (or V) tile
  static constexpr int kPBytes = (BN / 64) * 128 * 128;  // one P buffer
  static constexpr int kOffK = kQBytes;
  static constexpr int kOffV = kOffK + STAGES * kKVBytes;
  static constexpr int kOffP = kOffV + STAGES * kKVBytes;
  static constexpr int kOffBar = kOffP + 2 * kPBytes;
  static constexpr int kTotal = kOffBar + 256 + 1024;
  static constexpr uint32_t kTmemCols = (2 * BN + kDv) <= 256 ? 256 : 512;
  static constexpr uint32_t kOCol = 2 * BN;
};

template <typename T, int D, int BN, int STAGES>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
               const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
               const __grid_constant__ CUtensorMap tmV1, const AttnDev p) {
  pdl_wait();
  using CF = AttnCfg<D, BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CF::kOffBar);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // STAGES
  uint64_t* k_empty = k_full + STAGES;     // STAGES
  uint64_t* v_full = k_empty + STAGES;     // STAGES
  uint64_t* v_empty = v_full + STAGES;     // STAGES
  uint64_t* s_full = v_empty + STAGES;     // 2
  uint64_t* p_full = s_full + 2;           // 2
  uint64_t* pv_done = p_full + 2;          // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x;
  const int head = blockIdx.y;
  const int frame = p.frames - 1 - (int)blockIdx.z;   // cond (two-segment) frames are scheduled first
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
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&pv_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CF::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  pdl_launch();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, CF::kQBytes);
#pragma unroll
      for (int c = 0; c < CF::kChunks; ++c)
        tma_load_4d(smem + c * (128 * 128), &tmQ, q_full, c * 64, head, qt * 128, frame);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < ntiles; ++j) {
        const int seg = j / tiles_per_seg;
        const int kt = j - seg * tiles_per_seg;
        const CUtensorMap* mk = seg == 0 ? &tmK0 : &tmK1;
        const CUtensorMap* mv = seg == 0 ? &tmV0 : &tmV1;
        const int fr = seg == 0 ? frame : ref;
        mbar_wait(&k_empty[stage], phase ^ 1, 0x41);
        mbar_arrive_expect_tx(&k_full[stage], CF::kKVBytes);
#pragma unroll
        for (int c = 0; c < CF::kChunks; ++c)
          tma_load_4d(smem + CF::kOffK + stage * CF::kKVBytes + c * (BN * 128), mk, &k_full[stage],
                      c * 64, head, kt * BN, fr);
        mbar_wait(&v_empty[stage], phase ^ 1, 0x42);
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
    constexpr uint32_t idesc_pv = make_idesc_f16(128, CF::kDv, Cvt<T>::kFmt, 0, 1);  // B (=V) MN-major
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK = smem_u32(smem + CF::kOffK);
    const uint32_t sV = smem_u32(smem + CF::kOffV);
    const uint32_t sP = smem_u32(smem + CF::kOffP);

    auto issue_qk = [&](int stage, int sbuf) {
      const uint32_t kbase = sK + stage * CF::kKVBytes;
#pragma unroll
      for (int k = 0; k < CF::kKSteps; ++k) {
        const uint32_t off_q = (k >> 2) * (128 * 128) + (k & 3) * 32;
        const uint32_t off_k = (k >> 2) * (BN * 128) + (k & 3) * 32;
        umma_f16_ss(tmem_base + sbuf * BN, make_desc_sw128(sQ + off_q, 16, 1024),
                    make_desc_sw128(kbase + off_k, 16, 1024), idesc_qk, k != 0);
      }
    };

    mbar_wait(q_full, 0, 0x51);
    int kstage = 0, vstage = 0;
    uint32_t kphase = 0, vphase = 0;
    mbar_wait(&k_full[0], 0, 0x52);
    tc_fence_after();
    if (lane == 0) {
      issue_qk(0, 0);
      umma_commit(&k_empty[0]);
      umma_commit(&s_full[0]);
    }
    __syncwarp();
    if (++kstage == STAGES) { kstage = 0; kphase ^= 1; }

    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) {
        mbar_wait(&k_full[kstage], kphase, 0x53);
        tc_fence_after();
        if (lane == 0) {
          issue_qk(kstage, (j + 1) & 1);
          umma_commit(&k_empty[kstage]);
          umma_commit(&s_full[(j + 1) & 1]);
        }
        __syncwarp();
        if (++kstage == STAGES) { kstage = 0; kphase ^= 1; }
      }
      mbar_wait(&p_full[j & 1], (j >> 1) & 1, 0x54);
      mbar_wait(&v_full[vstage], vphase, 0x55);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t pbase = sP + (j & 1) * CF::kPBytes;
        const uint32_t vbase = sV + vstage * CF::kKVBytes;
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          const uint32_t off_p = (k >> 2) * (128 * 128) + (k & 3) * 32;
          // V tile: [d-chunk][key][64 d] -> MN-major B: 16 keys = 2 KB further, next d-chunk LBO away
          umma_f16_ss(tmem_base + CF::kOCol, make_desc_sw128(pbase + off_p, 16, 1024),
                      make_desc_sw128(vbase + k * 2048, BN * 128, 1024), idesc_pv, (j | k) != 0);
        }
        umma_commit(&v_empty[vstage]);
        umma_commit(&pv_done[j & 1]);
      }
      __syncwarp();
      if (++vstage == STAGES) { vstage = 0; vphase ^= 1; }
    }
  } else {
    // ============================ softmax warps ============================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
    uint8_t* sP = smem + CF::kOffP;
    float m_ref = -INFINITY;    // the max the stored P / O are currently relative to (log2 domain)
    float l_sum = 0.f;

    for (int j = 0; j < ntiles; ++j) {
      const int sbuf = j & 1;
      const int kt = j % tiles_per_seg;
      const int key0 = kt * BN;
      mbar_wait(&s_full[sbuf], (j >> 1) & 1, 0x61);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_addr + sbuf * BN;

      // ---- pass 1: row max ----
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_x32(s_addr + c * 32, r);
        tmem_ld_wait();
        if (key0 + c * 32 + 32 <= p.L) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (key0 + c * 32 + i < p.L) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      mx *= p.scale_log2;
      // lazy rescale: only move the reference max when some row of this warp grew by more than 2^8
      const bool need = (mx > m_ref + 8.0f);
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_ref, mx);
        if (j > 0) {
          // O may only be touched once P_{j-1} V_{j-1} has completed
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1, 0x62);
          tc_fence_after();
          const float f = exp2f(m_ref - m_new);
          const uint32_t o_addr = tmem_base + lane_addr + CF::kOCol;
#pragma unroll
          for (int c = 0; c < CF::kDv / 8; ++c) {
            uint32_t r[8];
            tmem_ld_x8(o_addr + c * 8, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
            tmem_st_x8(o_addr + c * 8, r);
          }
          tmem_st_wait();
          l_sum *= f;
        }
        m_ref = m_new;
      }

      // ---- pass 2: P = 2^(s*scale - m_ref), row sum, write P (swizzled, K-major over keys) ----
      if (j >= 2) mbar_wait(&pv_done[sbuf], ((j - 2) >> 1) & 1, 0x63);   // P buffer free again
      uint8_t* prow = sP + sbuf * CF::kPBytes + row * 128;
      float psum = 0.f;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_x32(s_addr + c * 32, r);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = fast_exp2(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_ref));
          if (key0 + c * 32 + i >= p.L) e = 0.f;
          pv[i] = e;
          psum += e;
        }
        uint8_t* pchunk = prow + (c >> 1) * (128 * 128);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 o4;
          o4.x = Cvt<T>::pack2(pv[u * 8 + 0], pv[u * 8 + 1]);
          o4.y = Cvt<T>::pack2(pv[u * 8 + 2], pv[u * 8 + 3]);
          o4.z = Cvt<T>::pack2(pv[u * 8 + 4], pv[u * 8 + 5]);
          o4.w = Cvt<T>::pack2(pv[u * 8 + 6], pv[u * 8 + 7]);
          const int unit = ((c & 1) * 4 + u) ^ (row & 7);
          *reinterpret_cast<uint4*>(pchunk + unit * 16) = o4;
        }
      }
      l_sum += psum;
      fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sbuf]);
    }

    // ---- epilogue: O / l -> global ----
    mbar_wait(&pv_done[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1, 0x64);
    tc_fence_after();
    const float inv = 1.0f / l_sum;
    const int qrow = qt * 128 + row;
    T* out = reinterpret_cast<T*>(p.O) + ((long long)frame * p.L + qrow) * p.ldo + head * D;
    const uint32_t o_addr = tmem_base + lane_addr + CF::kOCol;
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
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<CF::kTmemCols>(tmem_base);
  }
}

// (d, head, token, frame) tensor map over a [frames*L, ld] token matrix whose columns are heads*D wide
static int make_qkv_map(CUtensorMap* m, int dtype, const void* base, int D, int heads, int L, int frames,
                        long long ld, int box_rows) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)heads, (uint64_t)L, (uint64_t)frames};
  uint64_t str[3] = {(uint64_t)D * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * L};
  uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return make_tmap_16b(m, dtype, base, 4, dims, str, box);
}

template <typename T, int D, int BN, int STAGES>
static int launch_attn(const hb_attention_params* q, cudaStream_t stream) {
  using CF = AttnCfg<D, BN, STAGES>;
  static_assert(CF::kTotal <= 232448, "attention smem budget");
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
  auto kern = attn_tc_kernel<T, D, BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    HB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::kTotal));
    attr_set = true;
  }
  dim3 grid((q->L + 127) / 128, q->heads, q->frames);
  launch_kernel(kern, grid, kAttnThreads, CF::kTotal, stream, tmQ, tmK0, tmV0, tmK1, tmV1, d);
  HB_LAUNCH_CHECK();
  return HB_OK;
}

}  // namespace hb
