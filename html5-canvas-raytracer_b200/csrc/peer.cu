// peer.cu — framebuffer exchange between the GPUs of one box over peer memory (NVLink / NVSwitch), the
// multi-GPU mode of SURVEY.md §8e without a collective library in the data path.
//
// Every rank renders its interleaved row stripes (packed), then ONE kernel of its own stores those rows
// into the full-size framebuffer of EVERY rank at their final position — the all-gather and the
// de-interleave in one pass of plain 16-byte stores through the NVLink fabric — and publishes a sequence
// number in each peer's control block.  A frame is complete on a rank once all `world` sequence numbers
// have arrived; a one-warp wait kernel holds the consumer's stream until then.  Nobody spins while
// holding more than one warp, and no kernel of one rank waits for SM slots on another — that is what the
// persistent render kernels of several frames in flight cost an NCCL all-gather (its CTAs wait for peers
// while occupying SMs).  Buffers are shared between the per-GPU processes with CUDA IPC handles, which
// the host layer exchanges once over torch.distributed.
//
// Memory ordering: remote data stores -> __threadfence_system() -> device atomic (last-block election) ->
// __threadfence_system() -> remote flag store; the reader polls the flag (volatile), fences, then reads.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include "rtb_internal.h"

namespace {

struct PeerPtrs {
  void* p[RTB_MAX_PEERS];
};

// local packed rows -> every rank's frame, for the n_frames frames of a group (frame f: local rows at
// local + f * local_stride quads, destination frames.p[p] + f * frame_stride quads).  One thread moves
// 16 bytes (4 pixels) to all peers.
__global__ void __launch_bounds__(256)
peer_scatter_kernel(const uint4* __restrict__ local, PeerPtrs frames, PeerPtrs ctl, int quads_per_row, int n_local,
                    int n_frames, size_t local_stride, size_t frame_stride, RowMap rm, int world, int rank,
                    const uint32_t* __restrict__ status_word, uint32_t ready_off, uint32_t status_off, uint32_t seq,
                    uint32_t* counter) {
  const long long per_frame = (long long)n_local * quads_per_row;
  const long long total = per_frame * n_frames;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / per_frame);
    const long long k = i - (long long)f * per_frame;
    const int l = (int)(k / quads_per_row), q = (int)(k % quads_per_row);
    const uint4 v = local[(size_t)f * local_stride + (size_t)k];
    const size_t dst = (size_t)f * frame_stride + (size_t)rm.local_to_y(l) * quads_per_row + q;
#pragma unroll 4
    for (int p = 0; p < world; p++) static_cast<uint4*>(frames.p[p])[dst] = v;
  }
  // last block out publishes: status word first, then the sequence number
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < world) {
      uint32_t* c = static_cast<uint32_t*>(ctl.p[threadIdx.x]);
      c[status_off + rank] = status_word ? *status_word : 0u;
      __threadfence_system();
      *reinterpret_cast<volatile uint32_t*>(c + ready_off + rank) = seq;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

__global__ void peer_signal_kernel(PeerPtrs ctl, int world, int rank, uint32_t off, uint32_t value) {
  __threadfence_system();
  if ((int)threadIdx.x < world)
    *reinterpret_cast<volatile uint32_t*>(static_cast<uint32_t*>(ctl.p[threadIdx.x]) + off + rank) = value;
}

// Hold the stream until words[0..n) have all reached `value` (sequence numbers only grow; compared as
// signed distance so that wrap-around after 2^32 frames is harmless).  Gives up after timeout_ns and
// raises *timeout_flag instead of hanging the GPU if a peer has died.
__global__ void peer_wait_kernel(const uint32_t* words, int n, uint32_t value, unsigned long long timeout_ns,
                                 uint32_t* timeout_flag) {
  if ((int)threadIdx.x >= n) return;
  const volatile uint32_t* w = words + threadIdx.x;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned spins = 0;
  while ((int32_t)(*w - value) < 0) {
    if ((++spins & 0x3ffu) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) { *timeout_flag = 1u; break; }
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

// Stream memory operations of the driver API (cuStreamWaitValue32): the stream itself waits for a word
// in device memory — no kernel, no SM slot.  A spinning wait kernel is only one warp, but the persistent
// render kernels own every register of every SM, so it can only run in the place of one of their CTAs;
// with eight frames in flight up to sixteen such waits would be resident.  Looked up at run time so that
// the library keeps linking against the runtime API only.
typedef int (*StreamWaitValue32Fn)(void* stream, unsigned long long addr, uint32_t value, unsigned int flags);
StreamWaitValue32Fn stream_wait_fn() {
  static bool looked = false;
  static StreamWaitValue32Fn fn = nullptr;
  if (!looked) {
    looked = true;
    if (!getenv("RTB_PEER_KERNEL_WAIT")) {
      void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
      if (h) {
        fn = (StreamWaitValue32Fn)dlsym(h, "cuStreamWaitValue32_v2");
        if (!fn) fn = (StreamWaitValue32Fn)dlsym(h, "cuStreamWaitValue32");
      }
    }
  }
  return fn;
}

int check_peers(int world, int rank) {
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (world < 1 || world > RTB_MAX_PEERS || rank < 0 || rank >= world) {
    rtb_set_error("bad peer arguments: world %d (max %d), rank %d", world, RTB_MAX_PEERS, rank);
    return RTB_ERR_INVALID;
  }
  return RTB_OK;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int rtb_peer_alloc(uint64_t bytes, void** d_ptr, uint8_t* handle64) {
  RTB_API_LOCK();
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (!d_ptr || !handle64 || bytes == 0) { rtb_set_error("NULL argument"); return RTB_ERR_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == RTB_PEER_HANDLE_BYTES, "handle size");
  void* p = nullptr;
  RTB_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    rtb_set_error("peer buffer: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return RTB_ERR_CUDA;
  }
  memcpy(handle64, &h, sizeof h);
  *d_ptr = p;
  return RTB_OK;
}

int rtb_peer_open(const uint8_t* handle64, void** d_ptr) {
  RTB_API_LOCK();
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (!d_ptr || !handle64) { rtb_set_error("NULL argument"); return RTB_ERR_INVALID; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof h);
  RTB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RTB_OK;
}

int rtb_peer_close(void* d_ptr) {
  RTB_API_LOCK();
  if (!d_ptr) return RTB_OK;
  RTB_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return RTB_OK;
}

int rtb_peer_free(void* d_ptr) {
  RTB_API_LOCK();
  if (!d_ptr) return RTB_OK;
  RTB_CUDA(cudaFree(d_ptr));
  return RTB_OK;
}

int rtb_peer_scatter(const void* d_local_packed, void* const* peer_frames, void* const* peer_ctl, int32_t width,
                     int32_t height, int32_t stripe_rows, int32_t stripe_rank, int32_t stripe_world,
                     const void* d_status_u32, uint32_t ready_offset, uint32_t status_offset, uint32_t seq,
                     void* cuda_stream) {
  return rtb_peer_scatter_group(d_local_packed, 0, 1, peer_frames, 0, peer_ctl, width, height, stripe_rows,
                                stripe_rank, stripe_world, d_status_u32, ready_offset, status_offset, seq, cuda_stream);
}

int rtb_peer_scatter_group(const void* d_local_packed, uint64_t local_frame_stride_bytes, int32_t n_frames,
                           void* const* peer_frames, uint64_t peer_frame_stride_bytes, void* const* peer_ctl,
                           int32_t width, int32_t height, int32_t stripe_rows, int32_t stripe_rank,
                           int32_t stripe_world, const void* d_status_u32, uint32_t ready_offset,
                           uint32_t status_offset, uint32_t seq, void* cuda_stream) {
  RTB_API_LOCK();
  int rc = check_peers(stripe_world, stripe_rank);
  if (rc != RTB_OK) return rc;
  if (!d_local_packed || !peer_frames || !peer_ctl || width <= 0 || height <= 0 || stripe_rows <= 0 ||
      n_frames < 1 || n_frames > RTB_MAX_GROUP || (local_frame_stride_bytes & 15u) || (peer_frame_stride_bytes & 15u)) {
    rtb_set_error("bad scatter arguments");
    return RTB_ERR_INVALID;
  }
  if (width % 4 != 0) { rtb_set_error("peer scatter moves 16-byte quads: width must be a multiple of 4"); return RTB_ERR_UNSUPPORTED; }
  PeerPtrs fr, ct;
  memset(&fr, 0, sizeof fr); memset(&ct, 0, sizeof ct);
  for (int p = 0; p < stripe_world; p++) {
    if (!peer_frames[p] || !peer_ctl[p]) { rtb_set_error("peer %d: NULL buffer", p); return RTB_ERR_INVALID; }
    fr.p[p] = peer_frames[p]; ct.p[p] = peer_ctl[p];
  }
  if (!g_rtb.d_peer_counter) {
    RTB_CUDA(cudaMalloc(&g_rtb.d_peer_counter, RTB_MAX_IN_FLIGHT * sizeof(uint32_t)));
    RTB_CUDA(cudaMemset(g_rtb.d_peer_counter, 0, RTB_MAX_IN_FLIGHT * sizeof(uint32_t)));
  }
  RowMap rm;
  rm.y0 = 0; rm.y1 = height; rm.S = stripe_rows; rm.rank = stripe_rank; rm.world = stripe_world;
  rm.packed = 1; rm.width = width;
  const int n_local = rm.n_local();
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : g_rtb.stream;
  // scatters of different frames may overlap on different streams: one election counter per sequence slot
  uint32_t* counter = g_rtb.d_peer_counter + (seq % RTB_MAX_IN_FLIGHT);
  const int quads = width / 4;
  const long long total = (long long)n_local * quads * n_frames;
  int grid = (int)std::min<long long>((total + 255) / 256, (long long)g_rtb.sm_count * 4);
  if (grid < 1) grid = 1;
  peer_scatter_kernel<<<grid, 256, 0, st>>>(static_cast<const uint4*>(d_local_packed), fr, ct, quads, n_local, n_frames,
                                            (size_t)(local_frame_stride_bytes / 16), (size_t)(peer_frame_stride_bytes / 16),
                                            rm, stripe_world, stripe_rank, static_cast<const uint32_t*>(d_status_u32),
                                            ready_offset, status_offset, seq, counter);
  RTB_CUDA(cudaGetLastError());
  return RTB_OK;
}

int rtb_peer_signal(void* const* peer_ctl, int32_t world, int32_t rank, uint32_t word_offset, uint32_t value,
                    void* cuda_stream) {
  RTB_API_LOCK();
  int rc = check_peers(world, rank);
  if (rc != RTB_OK) return rc;
  if (!peer_ctl) { rtb_set_error("NULL argument"); return RTB_ERR_INVALID; }
  PeerPtrs ct;
  memset(&ct, 0, sizeof ct);
  for (int p = 0; p < world; p++) {
    if (!peer_ctl[p]) { rtb_set_error("peer %d: NULL buffer", p); return RTB_ERR_INVALID; }
    ct.p[p] = peer_ctl[p];
  }
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : g_rtb.stream;
  peer_signal_kernel<<<1, 32, 0, st>>>(ct, world, rank, word_offset, value);
  RTB_CUDA(cudaGetLastError());
  return RTB_OK;
}

int rtb_peer_wait(const void* d_words_u32, int32_t n, uint32_t value, uint32_t timeout_ms, void* d_timeout_flag_u32,
                  void* cuda_stream) {
  RTB_API_LOCK();
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (!d_words_u32 || !d_timeout_flag_u32 || n < 1 || n > 32) { rtb_set_error("bad wait arguments"); return RTB_ERR_INVALID; }
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : g_rtb.stream;
  if (timeout_ms == 0) {
    if (StreamWaitValue32Fn wait = stream_wait_fn()) {
      bool ok = true;
      for (int i = 0; i < n && ok; i++)                       // flags 0 = CU_STREAM_WAIT_VALUE_GEQ (cyclic compare)
        ok = wait((void*)st, (unsigned long long)(uintptr_t)(static_cast<const uint32_t*>(d_words_u32) + i), value, 0u) == 0;
      if (ok) return RTB_OK;
      rtb_set_error("cuStreamWaitValue32 failed");
      return RTB_ERR_CUDA;
    }
    timeout_ms = 20000;                                       // driver without stream memory operations
  }
  peer_wait_kernel<<<1, 32, 0, st>>>(static_cast<const uint32_t*>(d_words_u32), n, value,
                                     (unsigned long long)timeout_ms * 1000000ull,
                                     static_cast<uint32_t*>(d_timeout_flag_u32));
  RTB_CUDA(cudaGetLastError());
  return RTB_OK;
}

}  // extern "C"
#pragma GCC visibility pop
