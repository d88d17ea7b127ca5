// hallo_b200_cross_attention on tcgen05: attention of every query row against the <= 32 keys of its frame
// (4 image tokens: mutual_self_attention.py:289-303; 32 audio tokens x 3 mask regions: attention.py:854-890).
//
// One CTA = one (frame, head, region, query slab).  The K / V tiles (NK keys x d) are loaded once and stay in
// shared memory; the CTA then streams its 128-row query tiles through a 2-deep pipeline:
//   warp 0      TMA: K, V once; Q tiles into a 2-slot ring
//   warp 1      MMA: S[b] = Q_i K^T (M=128, N=NK), O[b] = P_i V (A = P in TMEM)
//   warps 2-5   softmax over the NK columns (one thread per row), P -> TMEM, then O[b] / l -> global
// S, P and O are double-buffered in TMEM so the MMAs of tile i+1 overlap the softmax / store of tile i.
// Keys beyond n_keys (image tokens: 4 of the 16-wide MMA tile) are masked to -inf.
#include <cstdlib>

#include "host_common.cuh"
#include "ptx.cuh"

namespace hb {

constexpr int kXThreads = 192;

struct XAttnDev {
  int L, heads, regions, n_keys, kv_frame_div, slabs, tiles_per_slab;
  void* O;
  long long ldo;
  int o_region_stride;
  float scale_log2;
};

template <int D, int NK>
struct XCfg {
  static constexpr int kChunks = (D + 63) / 64;
  static constexpr int kKSteps = (D + 15) / 16;
  static constexpr int kDv = ((D + 15) / 16) * 16;
  static constexpr int kQBytes = kChunks * 128 * 128;
  static constexpr int kKVRows = 64;                       // smem rows reserved per chunk (NK <= 32 used)
  static constexpr int kKVBytes = kChunks * kKVRows * 128;
  static constexpr int kOffK = 2 * kQBytes;
  static constexpr int kOffV = kOffK + kKVBytes;
  static constexpr int kOffBar = kOffV + kKVBytes;
  static constexpr int kTotal = kOffBar + 256 + 1024;
  static constexpr uint32_t kSStride = 32;                 // S[b] at b*32 (NK <= 32 columns; P aliases the first NK/2)
  static constexpr uint32_t kOCol0 = 64;
  static constexpr uint32_t kOStride = ((kDv + 31) / 32) * 32;
  static constexpr uint32_t kTmemCols = (kOCol0 + 2 * kOStride) <= 256 ? 256 : 512;
  static_assert(kOCol0 + 2 * kOStride <= 512, "TMEM budget");
};

template <typename T, int D, int NK>
__global__ void __launch_bounds__(kXThreads, 1)
xattn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const XAttnDev p) {
  pdl_wait();
  using CF = XCfg<D, NK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CF::kOffBar);
  uint64_t* kv_full = bars;            // 1
  uint64_t* q_full = bars + 1;         // 2
  uint64_t* q_empty = q_full + 2;      // 2
  uint64_t* s_full = q_empty + 2;      // 2
  uint64_t* p_full = s_full + 2;       // 2
  uint64_t* o_full = p_full + 2;       // 2
  uint64_t* o_empty = o_full + 2;      // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int slab = blockIdx.x;
  const int frame = blockIdx.y;
  const int head = blockIdx.z % p.heads;
  const int region = blockIdx.z / p.heads;
  const int hcol = region * p.heads + head;            // "head" index inside the region-major column layout of Q / O
  const int tiles_total = (p.L + 127) / 128;
  const int t0 = slab * p.tiles_per_slab;
  const int ntiles = min(p.tiles_per_slab, tiles_total - t0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 1);
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 4);
      mbar_init(&o_full[b], 1);
      mbar_init(&o_empty[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CF::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  pdl_launch();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (ntiles > 0) {
    if (warp == 0) {
      // ============================ TMA producer ============================
      if (lane == 0) {
        const int kvf = frame / p.kv_frame_div;
        // K / V column layout: region r, head h at "head" index (2r*heads + h) for K, ((2r+1)*heads + h) for V
        mbar_arrive_expect_tx(kv_full, 2 * CF::kChunks * NK * 128);
#pragma unroll
        for (int c = 0; c < CF::kChunks; ++c) {
          tma_load_4d(smem + CF::kOffK + c * (CF::kKVRows * 128), &tmK, kv_full, c * 64, 2 * region * p.heads + head, 0, kvf);
          tma_load_4d(smem + CF::kOffV + c * (CF::kKVRows * 128), &tmV, kv_full, c * 64, 2 * region * p.heads + head, 0, kvf);
        }
        for (int i = 0; i < ntiles; ++i) {
          const int b = i & 1;
          mbar_wait(&q_empty[b], ((i >> 1) & 1) ^ 1, 0xA1);
          mbar_arrive_expect_tx(&q_full[b], CF::kQBytes);
#pragma unroll
          for (int c = 0; c < CF::kChunks; ++c)
            tma_load_4d(smem + b * CF::kQBytes + c * (128 * 128), &tmQ, &q_full[b], c * 64, hcol, (t0 + i) * 128, frame);
        }
      }
    } else if (warp == 1) {
      // ============================ MMA issuer ============================
      constexpr uint32_t idesc_qk = make_idesc_f16(128, NK, Cvt<T>::kFmt, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, CF::kDv, Cvt<T>::kFmt, 0, 1);
      const uint32_t sQ = smem_u32(smem);
      const uint32_t sK = smem_u32(smem + CF::kOffK);
      const uint32_t sV = smem_u32(smem + CF::kOffV);
      auto issue_qk = [&](int b) {
#pragma unroll
        for (int k = 0; k < CF::kKSteps; ++k) {
          const uint32_t off_q = b * CF::kQBytes + (k >> 2) * (128 * 128) + (k & 3) * 32;
          const uint32_t off_k = (k >> 2) * (CF::kKVRows * 128) + (k & 3) * 32;
          umma_f16_ss(tmem_base + b * CF::kSStride, make_desc_sw128(sQ + off_q, 16, 1024),
                      make_desc_sw128(sK + off_k, 16, 1024), idesc_qk, k != 0);
        }
      };
      mbar_wait(kv_full, 0, 0xB1);
      mbar_wait(&q_full[0], 0, 0xB2);
      tc_fence_after();
      if (lane == 0) {
        issue_qk(0);
        umma_commit(&s_full[0]);
      }
      __syncwarp();
      for (int i = 0; i < ntiles; ++i) {
        const int b = i & 1;
        if (i + 1 < ntiles) {
          // S[b^1] / P[b^1] were consumed by P V of tile i-1, issued earlier in order
          mbar_wait(&q_full[b ^ 1], ((i + 1) >> 1) & 1, 0xB3);
          tc_fence_after();
          if (lane == 0) {
            issue_qk(b ^ 1);
            umma_commit(&s_full[b ^ 1]);
          }
          __syncwarp();
        }
        mbar_wait(&p_full[b], (i >> 1) & 1, 0xB4);
        mbar_wait(&o_empty[b], ((i >> 1) & 1) ^ 1, 0xB5);
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < NK / 16; ++k) {
            umma_f16_ts(tmem_base + CF::kOCol0 + b * CF::kOStride, tmem_base + b * CF::kSStride + k * 8,
                        make_desc_sw128(sV + k * 2048, CF::kKVRows * 128, 1024), idesc_pv, k != 0);
          }
          umma_commit(&o_full[b]);
          umma_commit(&q_empty[b]);
        }
        __syncwarp();
      }
    } else {
      // ============================ softmax + epilogue warps ============================
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
      for (int i = 0; i < ntiles; ++i) {
        const int b = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait(&s_full[b], ph, 0xC1);
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_addr + b * CF::kSStride;
        uint32_t s[NK];
        if (NK == 32) tmem_ld_x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        else tmem_ld_x16(s_addr, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
        tmem_ld_wait();
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < NK; ++k)
          if (k < p.n_keys) mx = fmaxf(mx, __uint_as_float(s[k]));
        mx *= p.scale_log2;
        float sum = 0.f;
        float e[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
          e[k] = (k < p.n_keys) ? fast_exp2(fmaf(__uint_as_float(s[k]), p.scale_log2, -mx)) : 0.f;
          sum += e[k];
        }
        uint32_t pk[NK / 2];
#pragma unroll
        for (int k = 0; k < NK; k += 2) pk[k >> 1] = Cvt<T>::pack2(e[k], e[k + 1]);
        if (NK == 32) tmem_st_x16(s_addr, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
        else tmem_st_x8(s_addr, *reinterpret_cast<uint32_t(*)[8]>(&pk[0]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);

        // ---- O[b] / l -> global ----
        mbar_wait(&o_full[b], ph, 0xC2);
        tc_fence_after();
        const float inv = 1.0f / sum;
        const int qrow = (t0 + i) * 128 + row;
        T* out = reinterpret_cast<T*>(p.O) + ((long long)frame * p.L + qrow) * p.ldo + region * p.o_region_stride + head * D;
        const uint32_t o_addr = tmem_base + lane_addr + CF::kOCol0 + b * CF::kOStride;
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
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[b]);
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

// (d, "head", token, frame) tensor map; "heads" counts every d-wide column group of the buffer
static int make_x_map(CUtensorMap* m, int dtype, const void* base, int D, int ncolgroups, int rows_per_frame,
                      int frames, long long ld, int box_rows) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)ncolgroups, (uint64_t)rows_per_frame, (uint64_t)frames};
  uint64_t str[3] = {(uint64_t)D * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * rows_per_frame};
  uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return make_tmap_16b(m, dtype, base, 4, dims, str, box);
}

template <typename T, int D, int NK>
static int launch_xattn(int dtype, const void* Q, long long ldq, const void* K, const void* V, long long ldkv,
                        void* O, long long ldo, int o_region_stride, int frames, int L, int heads, int n_keys,
                        int kv_frame_div, int regions, cudaStream_t stream) {
  using CF = XCfg<D, NK>;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  const int kv_frames = (frames + kv_frame_div - 1) / kv_frame_div;
  // Q columns: region r, head h at r*heads*D + h*D (q_region_stride == heads*D); K / V: [K_r | V_r] pairs of heads*D
  if ((rc = make_x_map(&tmQ, dtype, Q, D, regions * heads, L, frames, ldq, 128))) return rc;
  if ((rc = make_x_map(&tmK, dtype, K, D, 2 * regions * heads, n_keys, kv_frames, ldkv, NK))) return rc;
  if ((rc = make_x_map(&tmV, dtype, V, D, 2 * regions * heads, n_keys, kv_frames, ldkv, NK))) return rc;
  XAttnDev d{};
  d.L = L;
  d.heads = heads;
  d.regions = regions;
  d.n_keys = n_keys;
  d.kv_frame_div = kv_frame_div;
  d.O = O;
  d.ldo = ldo;
  d.o_region_stride = o_region_stride;
  d.scale_log2 = (float)(1.4426950408889634 / sqrt((double)D));
  const int tiles_total = (L + 127) / 128;
  // enough CTAs for ~4 waves, but at least 4 tiles per CTA so that the K/V load and pipeline fill amortise
  long long base_ctas = (long long)frames * heads * regions;
  int slabs = (int)((4LL * num_sms() + base_ctas - 1) / base_ctas);
  if (slabs < 1) slabs = 1;
  int tps = (tiles_total + slabs - 1) / slabs;
  if (tps < 4) tps = tiles_total < 4 ? tiles_total : 4;
  slabs = (tiles_total + tps - 1) / tps;
  d.slabs = slabs;
  d.tiles_per_slab = tps;
  auto kern = xattn_tc_kernel<T, D, NK>;
  static bool attr_set = false;
  if (!attr_set) {
    HB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::kTotal));
    attr_set = true;
  }
  dim3 grid(slabs, frames, heads * regions);
  launch_kernel(kern, grid, kXThreads, CF::kTotal, stream, tmQ, tmK, tmV, d);
  HB_LAUNCH_CHECK();
  return HB_OK;
}

template <typename T>
static int dispatch_xattn(int dtype, const void* Q, long long ldq, const void* K, const void* V, long long ldkv, void* O,
                          long long ldo, int o_region_stride, int frames, int L, int heads, int head_dim, int n_keys,
                          int kv_frame_div, int regions, cudaStream_t s) {
#define HB_X(D_, NK_) \
  return launch_xattn<T, D_, NK_>(dtype, Q, ldq, K, V, ldkv, O, ldo, o_region_stride, frames, L, heads, n_keys, kv_frame_div, regions, s)
  if (n_keys == 32) {
    if (head_dim == 40) HB_X(40, 32);
    if (head_dim == 80) HB_X(80, 32);
    if (head_dim == 160) HB_X(160, 32);
  } else {
    if (head_dim == 40) HB_X(40, 16);
    if (head_dim == 80) HB_X(80, 16);
    if (head_dim == 160) HB_X(160, 16);
  }
#undef HB_X
  return fail(HB_ERR_BAD_SHAPE, "cross_attention (tcgen05): head_dim %d", head_dim);
}

// returns HB_OK if handled, 1 if the shape is left to the CUDA-core kernel
int xattn_tc_try(int dtype, const void* Q, long long ldq, int q_region_stride, const void* K, const void* V,
                 long long ldkv, int kv_region_stride, void* O, long long ldo, int o_region_stride, int frames, int L,
                 int heads, int head_dim, int n_keys, int kv_frame_div, int regions, cudaStream_t s) {
  // Opt-in until it has been parity-tested on hardware (written after the round-1 GPU budget was spent):
  // xattn_tc = 1 (HALLO_B200_XATTN_TC) routes eligible shapes here; default is the tested CUDA-core kernel in aux.cu.
  const bool off = option(OPT_XATTN_TC) == 0;
  const int C = heads * head_dim;
  // layout contract of this kernel: Q regions C apart, [K_r | V_r] pairs 2C apart with V = K + C columns, n_keys <= 32
  if (off || (head_dim != 40 && head_dim != 80 && head_dim != 160) || n_keys > 32 || n_keys < 1 || L < 128) return 1;
  if (regions > 1 && (q_region_stride != C || kv_region_stride != 2 * C)) return 1;
  if (reinterpret_cast<const char*>(V) - reinterpret_cast<const char*>(K) != (long long)C * 2) return 1;
  if (dtype == HB_F16)
    return dispatch_xattn<__half>(dtype, Q, ldq, K, V, ldkv, O, ldo, o_region_stride, frames, L, heads, head_dim, n_keys,
                                  kv_frame_div, regions, s);
  if (dtype == HB_BF16)
    return dispatch_xattn<__nv_bfloat16>(dtype, Q, ldq, K, V, ldkv, O, ldo, o_region_stride, frames, L, heads, head_dim,
                                         n_keys, kv_frame_div, regions, s);
  return 1;
}

}  // namespace hb
