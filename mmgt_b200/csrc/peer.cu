// Peer memory over NVLink for the frame-shard <-> token-shard exchange around the motion modules
// (SURVEY.md section 8e level 3; motion_module.py:361-363,386).
//
//   * receive buffers are plain cudaMalloc allocations exported / imported with CUDA IPC (one process per GPU);
//   * the exchange itself is written by the producing kernel (gemm_tc.cu epilogue) or by row_exchange_kernel below:
//     every destination row is stored straight into the owning shard's buffer;
//   * peer_barrier_kernel is the only synchronisation: release-store of an epoch into every peer's flag array,
//     acquire-spin on the own array.  Two receive buffers alternate, so the barrier of exchange n+1 also proves
//     that every peer is done reading what exchange n delivered (see mmgt_b200/frame_shard.py).
#include <string.h>

#include "common.cuh"

int mmgt_row_exchange_check(const mmgt_row_exchange* ex, int64_t rows, const char* who) {
  MMGT_CHECK_ARG(ex != nullptr, MMGT_E_INVALID, "%s: null exchange", who);
  MMGT_CHECK_ARG(ex->k >= 1 && ex->k <= MMGT_MAX_PEERS && ex->my >= 0 && ex->my < ex->k, MMGT_E_INVALID,
                 "%s: bad shard %d of %d", who, ex->my, ex->k);
  MMGT_CHECK_ARG(ex->direction == 1 || ex->direction == 2, MMGT_E_INVALID, "%s: direction must be 1 or 2", who);
  MMGT_CHECK_ARG(ex->B > 0 && ex->F > 0 && ex->T > 0 && ex->F % ex->k == 0 && ex->T % ex->k == 0, MMGT_E_INVALID,
                 "%s: B=%d F=%d T=%d must be positive and F, T divisible by k=%d", who, ex->B, ex->F, ex->T, ex->k);
  MMGT_CHECK_ARG(rows == (int64_t)ex->B * (ex->F / ex->k) * ex->T, MMGT_E_INVALID,
                 "%s: %lld source rows, exchange expects B*(F/k)*T = %lld", who, (long long)rows,
                 (long long)ex->B * (ex->F / ex->k) * ex->T);
  for (int s = 0; s < ex->k; ++s)
    MMGT_CHECK_ARG(ex->peer_base[s] != nullptr && aligned16(ex->peer_base[s]), MMGT_E_ALIGN,
                   "%s: peer_base[%d] null or not 16-byte aligned", who, s);
  MMGT_CHECK_ARG(ex->ld > 0, MMGT_E_INVALID, "%s: ld must be positive", who);
  return 0;
}

namespace {

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void peer_barrier_kernel(mmgt_peer_barrier_params p) {
  const int lane = threadIdx.x;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(p.epoch) + 1u;
  __syncwarp();
  const bool peer = lane < p.k && lane != p.my;
  const bool dead = *reinterpret_cast<volatile uint32_t*>(p.status) != 0u;   // an earlier barrier timed out: do not wait again
  if (peer) {
    __threadfence_system();   // every store of the earlier kernels on this stream is ordered before the flag
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flags[lane] + p.my), "r"(e) : "memory");
  }
  if (peer && !dead) {
    const uint32_t* mine = p.flags[p.my] + lane;
    const uint64_t t0 = globaltimer_ns();
    const uint64_t limit = (uint64_t)(p.timeout_ms > 0 ? p.timeout_ms : 2000) * 1000000ull;
    uint32_t v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - e) >= 0) break;
      if (globaltimer_ns() - t0 > limit) {   // a peer died or ran a different schedule: report, never hang the GPU
        atomicExch(p.status, 1u);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncwarp();
  if (lane == 0) *reinterpret_cast<volatile uint32_t*>(p.epoch) = e;
}

// one warp per source row, 16-byte units
__global__ void row_exchange_kernel(const uint8_t* __restrict__ src, int64_t src_row_bytes, int row_bytes, int esize,
                                    int64_t rows, mmgt_row_exchange ex) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp0; m < rows; m += nwarps) {
    uint8_t* dst = esize == 2 ? reinterpret_cast<uint8_t*>(exchange_row_ptr<bf16>(ex, (int)m))
                              : reinterpret_cast<uint8_t*>(exchange_row_ptr<float>(ex, (int)m));
    const uint4* s4 = reinterpret_cast<const uint4*>(src + m * src_row_bytes);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int u = lane; u < row_bytes / 16; u += 32) d4[u] = s4[u];
  }
}

}  // namespace

extern "C" int mmgt_peer_alloc(mmgt_ctx* ctx, int64_t bytes, void** out_ptr) {
  MMGT_CHECK_ARG(ctx && out_ptr && bytes > 0, MMGT_E_INVALID, "peer_alloc: bad args");
  MMGT_CUDA_OK(cudaSetDevice(ctx->device));
  void* p = nullptr;
  MMGT_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  MMGT_CUDA_OK(cudaMemset(p, 0, (size_t)bytes));
  MMGT_CUDA_OK(cudaDeviceSynchronize());
  *out_ptr = p;
  return 0;
}

extern "C" int mmgt_peer_free(mmgt_ctx* ctx, void* ptr) {
  MMGT_CHECK_ARG(ctx != nullptr, MMGT_E_INVALID, "peer_free: null ctx");
  if (ptr) MMGT_CUDA_OK(cudaFree(ptr));
  return 0;
}

extern "C" int mmgt_peer_export(mmgt_ctx* ctx, const void* ptr, unsigned char* handle64_host) {
  MMGT_CHECK_ARG(ctx && ptr && handle64_host, MMGT_E_INVALID, "peer_export: bad args");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  MMGT_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64_host, &h, 64);
  return 0;
}

extern "C" int mmgt_peer_import(mmgt_ctx* ctx, const unsigned char* handle64_host, void** out_ptr) {
  MMGT_CHECK_ARG(ctx && handle64_host && out_ptr, MMGT_E_INVALID, "peer_import: bad args");
  MMGT_CUDA_OK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  void* p = nullptr;
  MMGT_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out_ptr = p;
  return 0;
}

extern "C" int mmgt_peer_unmap(mmgt_ctx* ctx, void* ptr) {
  MMGT_CHECK_ARG(ctx != nullptr, MMGT_E_INVALID, "peer_unmap: null ctx");
  if (ptr) MMGT_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int mmgt_peer_barrier(mmgt_ctx* ctx, const mmgt_peer_barrier_params* p, void* stream) {
  MMGT_CHECK_ARG(ctx && p, MMGT_E_INVALID, "peer_barrier: null ctx/params");
  MMGT_CHECK_ARG(p->k >= 1 && p->k <= MMGT_MAX_PEERS && p->my >= 0 && p->my < p->k && p->epoch && p->status, MMGT_E_INVALID,
                 "peer_barrier: bad shard %d of %d or null epoch/status", p->my, p->k);
  for (int s = 0; s < p->k; ++s) MMGT_CHECK_ARG(p->flags[s] != nullptr, MMGT_E_INVALID, "peer_barrier: flags[%d] is null", s);
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*p);
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int mmgt_row_exchange_copy(mmgt_ctx* ctx, const void* src, int64_t lds, int C, int dtype,
                                      const mmgt_row_exchange* ex, void* stream) {
  MMGT_CHECK_ARG(ctx && src && ex, MMGT_E_INVALID, "row_exchange_copy: null argument");
  MMGT_CHECK_ARG(dtype == MMGT_F32 || dtype == MMGT_BF16, MMGT_E_INVALID, "row_exchange_copy: bad dtype %d", dtype);
  const int esize = dtype == MMGT_F32 ? 4 : 2;
  const int64_t rows = (int64_t)ex->B * (ex->k > 0 ? ex->F / ex->k : 0) * ex->T;
  int rc = mmgt_row_exchange_check(ex, rows, "row_exchange_copy");
  if (rc) return rc;
  MMGT_CHECK_ARG(C > 0 && lds >= C && ex->ld >= C, MMGT_E_INVALID, "row_exchange_copy: leading dims smaller than C");
  MMGT_CHECK_ARG((C * esize) % 16 == 0 && (lds * esize) % 16 == 0 && (ex->ld * esize) % 16 == 0 && aligned16(src), MMGT_E_ALIGN,
                 "row_exchange_copy: rows must be 16-byte multiples and aligned");
  const int threads = 256;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
  row_exchange_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>((const uint8_t*)src, lds * esize, C * esize, esize, rows,
                                                                        *ex);
  MMGT_LAUNCH_OK(ctx);
  return 0;
}
