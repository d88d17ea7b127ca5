// C entry point of hallo_b200_attention: picks the kernel generation per shape.
#include "host_common.cuh"
#include <cstdlib>

namespace hb {

template <typename T>
static int dispatch_attn(const hb_attention_params* p, cudaStream_t s) {
  const bool v1 = option(OPT_ATTN_V1) != 0;     // A/B switches (hallo_b200_set_option / HALLO_B200_ATTN_*)
  const int poly = option(OPT_ATTN_POLY);
  const int occ2 = option(OPT_ATTN_OCC2);        // head_dim 40: 64-key steps, two CTAs (four query tiles) per SM
  switch (p->head_dim) {
    // v2 (two query tiles per CTA, P in TMEM) when a frame has at least one full pair of tiles; the
    // single-tile kernel otherwise (small L) and for head_dim 160 at small L.
    case 40:
      if (p->L >= 256 && !v1 && occ2 != 0) {
        if (poly == 4) return launch_attn2<T, 40, 64, 4, 2>(p, s);
        if (poly == 3) return launch_attn2<T, 40, 64, 3, 2>(p, s);
        if (poly == 2) return launch_attn2<T, 40, 64, 2, 2>(p, s);
        return launch_attn2<T, 40, 64, 0, 2>(p, s);
      }
      if (p->L >= 256 && !v1) {
        if (poly == 4) return launch_attn2<T, 40, 128, 4>(p, s);
        if (poly == 3) return launch_attn2<T, 40, 128, 3>(p, s);
        if (poly == 2) return launch_attn2<T, 40, 128, 2>(p, s);
        return launch_attn2<T, 40, 128, 0>(p, s);
      }
      return launch_attn<T, 40, 128, 2>(p, s);
    case 80:
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
