// Peer memory for the frame-sharded window (one process per GPU): CUDA-IPC exchange buffers, the flag barrier, and
// the residual add that closes a motion module.  The data movement itself is fused into the producing kernels
// (gn_apply_kernel's scattered store in aux.cu, the row-scatter epilogue of gemm_tc.cu); nothing here copies tensors.
#include "host_common.cuh"
#include "ptx.cuh"

namespace hb {

struct PeerFlags {
  unsigned int* flags[HB_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One CTA.  Thread t < n tells rank t "rank `me` has arrived at barrier e" (slot [me] of t's flag array) and waits for
// rank t's arrival in its own slot [t].  The epoch lives in local device memory so that graph replays keep counting.
__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerFlags pf, int n, int me, unsigned int* epoch) {
  __shared__ unsigned int ep;
  if (threadIdx.x == 0) {
    ep = *epoch + 1u;
    *epoch = ep;
  }
  __syncthreads();
  const unsigned int e = ep;
  __threadfence_system();                      // peer-memory stores of the preceding kernels on this stream, then the flag
  if ((int)threadIdx.x < n) {
    st_release_sys(pf.flags[threadIdx.x] + me, e);
    const unsigned int* mine = pf.flags[me] + threadIdx.x;
    const long long t0 = clock64();
    // signed distance: correct across the (theoretical) wrap of the 32-bit epoch
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > 40000000000LL) {    // ~20 s at 1.9 GHz: a peer died or fell out of step -- do not hang the GPU
        atomicExch(&g_hb_error, 0x70u | ((unsigned int)threadIdx.x << 8));
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  __threadfence_system();
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long nvec) {
  pdl_wait();
  pdl_launch();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 ua = reinterpret_cast<const uint4*>(a)[i];
    const uint4 ub = reinterpret_cast<const uint4*>(b)[i];
    const uint32_t* wa = reinterpret_cast<const uint32_t*>(&ua);
    const uint32_t* wb = reinterpret_cast<const uint32_t*>(&ub);
    uint4 uo;
    uint32_t* wo = reinterpret_cast<uint32_t*>(&uo);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = Cvt<T>::unpack2(wa[k]), y = Cvt<T>::unpack2(wb[k]);
      wo[k] = Cvt<T>::pack2(x.x + y.x, x.y + y.y);
    }
    reinterpret_cast<uint4*>(out)[i] = uo;
  }
}

}  // namespace hb

extern "C" int hallo_b200_peer_alloc(int64_t bytes, void** ptr_out) {
  if (!ptr_out || bytes <= 0) return hb::fail(HB_ERR_NULL, "peer_alloc: bad arguments");
  void* p = nullptr;
  HB_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
  HB_CUDA_CHECK(cudaMemset(p, 0, (size_t)bytes));
  HB_CUDA_CHECK(cudaDeviceSynchronize());
  *ptr_out = p;
  return HB_OK;
}

extern "C" int hallo_b200_peer_free(void* ptr) {
  if (ptr) HB_CUDA_CHECK(cudaFree(ptr));
  return HB_OK;
}

extern "C" int hallo_b200_peer_export(void* ptr, void* handle_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == HB_IPC_HANDLE_BYTES, "IPC handle size");
  if (!ptr || !handle_out) return hb::fail(HB_ERR_NULL, "peer_export: null pointer");
  cudaIpcMemHandle_t h;
  HB_CUDA_CHECK(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle_out, &h, sizeof(h));
  return HB_OK;
}

extern "C" int hallo_b200_peer_open(const void* handle, void** ptr_out) {
  if (!handle || !ptr_out) return hb::fail(HB_ERR_NULL, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  HB_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return HB_OK;
}

extern "C" int hallo_b200_peer_close(void* ptr) {
  if (ptr) HB_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return HB_OK;
}

extern "C" int hallo_b200_peer_barrier(void* const* flags, int n, int me, uint32_t* epoch, hb_stream_t stream) {
  if (!flags || !epoch) return hb::fail(HB_ERR_NULL, "peer_barrier: null pointer");
  if (n < 1 || n > HB_MAX_PEERS || me < 0 || me >= n) return hb::fail(HB_ERR_BAD_SHAPE, "peer_barrier: n=%d me=%d", n, me);
  hb::PeerFlags pf{};
  for (int i = 0; i < n; ++i) {
    if (!flags[i]) return hb::fail(HB_ERR_NULL, "peer_barrier: null flag array %d", i);
    pf.flags[i] = reinterpret_cast<unsigned int*>(flags[i]);
  }
  hb::peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pf, n, me, epoch);
  HB_LAUNCH_CHECK();
  return HB_OK;
}

extern "C" int hallo_b200_add(int dtype, const void* a, const void* b, void* out, int64_t n, hb_stream_t stream) {
  using namespace hb;
  if (!a || !b || !out) return fail(HB_ERR_NULL, "add: null pointer");
  if (n % 8 != 0 || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15))
    return fail(HB_ERR_BAD_SHAPE, "add: n=%lld must be a multiple of 8 and the pointers 16-byte aligned", (long long)n);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long nvec = n / 8;
  long long g = (nvec + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  const int grid = (int)(g < cap ? (g > 0 ? g : 1) : cap);
  if (dtype == HB_F16) launch_kernel(add_kernel<__half>, grid, 256, 0, s, (const __half*)a, (const __half*)b, (__half*)out, nvec);
  else if (dtype == HB_BF16)
    launch_kernel(add_kernel<__nv_bfloat16>, grid, 256, 0, s, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, nvec);
  else return fail(HB_ERR_BAD_DTYPE, "add: dtype %d", dtype);
  HB_LAUNCH_CHECK();
  return HB_OK;
}
