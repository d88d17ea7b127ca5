// C entry point of hallo_b200_attention: picks the kernel generation per shape.
#include "host_common.cuh"
#include <cstdlib>

namespace hb {

template <typename T>
static int dispatch_attn(const hb_attention_params* p, cudaStream_t s) {
  const bool v1 = option(OPT_ATTN_V1) != 0;     // A/B switches (hallo_b200_set_option / HALLO_B200_ATTN_*)
  const int poly = option(OPT_ATTN_POLY);
  // streamed softmax (attn2_tc.cu, CHUNK = 32): written after the last GPU session of round 1, so it stays
  // opt-in until tests/test_attention_gpu.py has passed with attn_chunk = 1 on hardware
  const int chunk_opt = option(OPT_ATTN_CHUNK);   // 1: streamed softmax, 2: + row sums on the tensor core (d = 40)
  const bool chunk = chunk_opt != 0;
  switch (p->head_dim) {
    // v2 (two query tiles per CTA, P in TMEM) when a frame has at least one full pair of tiles; the
    // single-tile kernel otherwise (small L) and for head_dim 160 at small L.
    case 40:
      if (p->L >= 256 && !v1 && option(OPT_ATTN_V3) != 0) {     // register-S kernel (attn3_tc.cu), opt-in
        if (poly == 4) return launch_attn3<T, 4>(p, s);
        if (poly == 3) return launch_attn3<T, 3>(p, s);
        if (poly == 2) return launch_attn3<T, 2>(p, s);
        return launch_attn3<T, 0>(p, s);
      }
      if (p->L >= 256 && !v1 && chunk_opt == 2) {
        if (poly == 4) return launch_attn2<T, 40, 128, 4, 32, true>(p, s);
        if (poly == 3) return launch_attn2<T, 40, 128, 3, 32, true>(p, s);
        return launch_attn2<T, 40, 128, 0, 32, true>(p, s);
      }
      if (p->L >= 256 && !v1 && chunk) {
        if (poly == 4) return launch_attn2<T, 40, 128, 4, 32>(p, s);
        if (poly == 3) return launch_attn2<T, 40, 128, 3, 32>(p, s);
        return launch_attn2<T, 40, 128, 0, 32>(p, s);
      }
      if (p->L >= 256 && !v1) {
        if (poly == 4) return launch_attn2<T, 40, 128, 4>(p, s);
        if (poly == 3) return launch_attn2<T, 40, 128, 3>(p, s);
        if (poly == 2) return launch_attn2<T, 40, 128, 2>(p, s);
        return launch_attn2<T, 40, 128, 0>(p, s);
      }
      return launch_attn<T, 40, 128, 2>(p, s);
    case 80:
      if (p->L >= 256 && !v1 && chunk) return launch_attn2<T, 80, 128, 0, 32>(p, s);
      if (p->L >= 256 && !v1) return launch_attn2<T, 80, 128, 0>(p, s);
      return launch_attn<T, 80, 64, 2>(p, s);
    case 160: return launch_attn<T, 160, 64, 2>(p, s);   // 2 x (2 S buffers + O) does not fit TMEM at d = 160
    default: return fail(HB_ERR_BAD_SHAPE, "attention: head_dim %d not in {40, 80, 160}", p->head_dim);
  }
}

}  // namespace hb

extern "C" int hallo_b200_attention(const hb_attention_params* p, hb_stream_t stream) {
  using namespace hb;
  if (p == nullptr || p->Q == nullptr || p->K == nullptr || p->V == nullptr || p->O == nullptr)
    return fail(HB_ERR_NULL, "hallo_b200_attention: null pointer");
  if (p->L <= 0 || p->frames <= 0 || p->heads <= 0 || p->heads > 256)
    return fail(HB_ERR_BAD_SHAPE, "hallo_b200_attention: L=%d frames=%d heads=%d", p->L, p->frames, p->heads);
  if (p->ldq % 8 || p->ldk % 8 || p->ldv % 8 || p->ldo % 8)
    return fail(HB_ERR_BAD_SHAPE, "hallo_b200_attention: leading dims must be multiples of 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (p->dtype == HB_F16) return dispatch_attn<__half>(p, s);
  if (p->dtype == HB_BF16) return dispatch_attn<__nv_bfloat16>(p, s);
  return fail(HB_ERR_BAD_DTYPE, "hallo_b200_attention: dtype %d", p->dtype);
}
