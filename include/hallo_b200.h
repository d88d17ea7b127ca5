/*
 * hallo_b200 -- C ABI of the sm_100a kernels behind the Hallo denoising hot path.
 *
 * The reference (fudan-generative-vision/hallo) has no FFI/operator layer of its own: its
 * hot path is Python classes over torch library calls.  This header is
 * therefore the boundary the build introduces (SURVEY.md 8b "B2"); every entry point names the
 * reference code whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless noted;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library never allocates
 *     or frees device memory;
 *   - all activations are "channels-last token matrices": row = (frame, pixel), column = channel;
 *   - every call is asynchronous on `stream` and re-entrant; returns 0 on success or a
 *     negative hb_status; never throws;
 *   - dtype selects the storage / tensor-core input type (fp32 accumulate everywhere).
 */
#ifndef HALLO_B200_H_
#define HALLO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* hb_stream_t; /* cudaStream_t */

enum hb_status {
  HB_OK = 0,
  HB_ERR_BAD_SHAPE = -1,   /* dimension not supported by the kernel (alignment, head dim ...) */
  HB_ERR_BAD_DTYPE = -2,
  HB_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed; see hallo_b200_last_error() */
  HB_ERR_NULL = -4,
  HB_ERR_DEVICE_TRAP = -5  /* a kernel recorded a barrier timeout in the device error word */
};

enum hb_dtype { HB_F16 = 0, HB_BF16 = 1 };

/* ------------------------------------------------------------------------------------------
 * Library info
 * ---------------------------------------------------------------------------------------- */
int hallo_b200_abi_version(void);
const char* hallo_b200_last_error(void);
/* Reads and clears the device-side error word written by a kernel that timed out on a barrier. */
int hallo_b200_device_error(unsigned int* code_out);
/* Number of kernel launches issued by this library since the last reset (bench.py gpu_launches). */
int64_t hallo_b200_launch_count(int reset);

/* ------------------------------------------------------------------------------------------
 * hallo_b200_gemm -- C = epilogue(A * W^T) on tcgen05 tensor cores, TMA-fed, TMEM accumulators.
 *
 * Replaces every nn.Linear / 1x1 conv / 3x3 conv on the path:
 *   to_q/to_k/to_v/to_out           diffusers Attention used at hallo/models/attention.py:479-503,
 *                                   703-761, hallo/models/motion_module.py:464-482
 *   FeedForward (GEGLU)             hallo/models/attention.py:517,777; motion_module.py:383
 *   proj_in / proj_out (1x1)        hallo/models/transformer_3d.py:197-203,236-251;
 *                                   hallo/models/motion_module.py:290-313
 *   zero_conv_{full,face,lip}       hallo/models/attention.py:854-890
 *   InflatedConv3d 3x3 (+shortcut)  hallo/models/resnet.py:50-66, 385-410, 166-183, 250
 *
 *   A  : [M, K] row-major (lda), optionally split along K into two sources (A for k < K1,
 *        A2 for k >= K1) -- the UNet skip-connection channel concat without a copy
 *        (hallo/models/unet_3d_blocks.py:1131,1373).
 *   W  : [N, K] row-major (torch Linear layout; conv weights packed [Cout][tap][Cin]).
 *   conv3x3 != 0: A is an NHWC image batch [img_n, img_h, img_w, Cin]; M = img_n*img_h*img_w,
 *        K = 9*Cin; stride 1, zero padding 1 (TMA out-of-bounds fill).
 *   epilogue, in this order (each optional):
 *        v  = acc + bias[col] + group_bias[row / rows_per_group][col]
 *        v  = v_even * gelu_erf(v_odd)           (HB_EPI_GEGLU: W rows interleaved value/gate,
 *                                                 output has N/2 columns)
 *        v  = v * row_scale[row] * alpha + residual[row][col]
 *   Constraints: K % 64 == 0 (K1 % 64 == 0, Cin % 64 == 0); lda/ldw/ldc/ldr % 8 == 0.
 * ---------------------------------------------------------------------------------------- */
enum hb_epi_flags { HB_EPI_GEGLU = 1 };

typedef struct {
  int32_t dtype;
  int32_t M, N, K;
  const void* A;
  int64_t lda;
  const void* A2; /* NULL when unused */
  int64_t lda2;
  int32_t K1;
  const void* W;
  int64_t ldw;
  void* C;
  int64_t ldc;
  const void* bias;       /* [N] */
  const void* group_bias; /* [ceil(M/rows_per_group), ld_group_bias] */
  int64_t ld_group_bias;
  int32_t rows_per_group;
  const void* row_scale; /* [M] */
  const void* residual;  /* [M, ldr] */
  int64_t ldr;
  float alpha;
  int32_t flags;
  int32_t conv3x3;
  int32_t img_n, img_h, img_w;
} hb_gemm_params;

int hallo_b200_gemm(const hb_gemm_params* p, hb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HALLO_B200_H_ */
