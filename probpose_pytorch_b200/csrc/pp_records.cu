// Decoded-keypoint records and their exchange between GPUs.
//
//  * pp_pack_records : the tail of Codec.decode (codec.py:249-263) on the device -- one packed (N, 7) float64 record
//                      (x, y, score, probability, visibility, oks, error / diagonal) per keypoint, the unit that is
//                      gathered across GPUs -- written in ONE kernel, and, when a mailbox is given, to
//                      every peer GPU in the same kernel: the final keypoint gather is plain stores into peer memory
//                      over NVLink (each rank writes its block into every rank's mailbox).  No collective call, no
//                      intermediate copy; latency-bound (a few hundred KB per rank).
//  * pp_mailbox_commit: stores the step's local loss next to the records and raises the per-source flag of the slot
//                      (system-scope release); launched after the records (and after the loss exists).
//  * pp_mailbox_wait : the consumer side -- waits (bounded) until every source rank's flag of a slot has reached the
//                      expected sequence number.
//
// Mailbox layout (identical on every rank, symmetric allocation): slots x world blocks of block_bytes
//   [ N * 7 doubles | pad to 16 | loss (double) | flag (uint32) | pad ]   with block_bytes = round_up(N * 56, 16) + 16.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/probpose_b200.h"
#include "pp_common.cuh"

namespace {

constexpr int kRecThreads = 256;

__host__ __device__ inline int64_t loss_offset(int64_t n_records) { return (n_records * 56 + 15) / 16 * 16; }

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kRecThreads)
pack_records_kernel(int64_t N, const double* __restrict__ keypoints, const float* __restrict__ scores,
                    const float* __restrict__ prob, const float* __restrict__ vis, const float* __restrict__ oks,
                    const float* __restrict__ err, float inv_diag, double* __restrict__ rec, pp_mailbox mb) {
  // the block's records are assembled in shared memory and leave as contiguous 16-byte stores: whole lines to the
  // local output and -- over NVLink -- to every rank's mailbox, instead of 8-byte stores 56 bytes apart
  __shared__ __align__(16) double tile[kRecThreads * 7];
  const int64_t first = static_cast<int64_t>(blockIdx.x) * kRecThreads;
  const int64_t n = first + threadIdx.x;
  if (n < N) {
    double* r = tile + threadIdx.x * 7;
    r[0] = keypoints[n * 2];
    r[1] = keypoints[n * 2 + 1];
    r[2] = static_cast<double>(scores[n]);
    r[3] = static_cast<double>(prob[n]);
    r[4] = static_cast<double>(vis[n]);
    r[5] = static_cast<double>(oks[n]);
    r[6] = static_cast<double>(__fmul_rn(err[n], inv_diag));   // float32 errors / float32 scalar, as torch does it
  }
  __syncthreads();
  const int count = static_cast<int>(min(static_cast<int64_t>(kRecThreads), N - first)) * 7;   // doubles in this block
  const int64_t base = first * 7;                                                              // even: 16-byte aligned
  const int pairs = count >> 1;
  const double2* src2 = reinterpret_cast<const double2*>(tile);
  if (rec) {
    double2* dst2 = reinterpret_cast<double2*>(rec + base);
    for (int i = threadIdx.x; i < pairs; i += kRecThreads) dst2[i] = src2[i];
    if ((count & 1) && threadIdx.x == 0) rec[base + count - 1] = tile[count - 1];
  }
  if (!mb.peer_bufs) return;
  const int64_t off = (static_cast<int64_t>(mb.slot) * mb.world + mb.rank) * mb.block_bytes;
  for (int p = 0; p < mb.world; ++p) {   // NVLink stores into every rank's mailbox (own rank included)
    double* dst = reinterpret_cast<double*>(static_cast<unsigned char*>(mb.peer_bufs[p]) + off) + base;
    double2* dst2 = reinterpret_cast<double2*>(dst);
    for (int i = threadIdx.x; i < pairs; i += kRecThreads) dst2[i] = src2[i];
    if ((count & 1) && threadIdx.x == 0) dst[count - 1] = tile[count - 1];
  }
  // one system-scope fence per block, after the block barrier (cumulative over the stores the barrier has ordered):
  // when this kernel has completed, its records are visible to every rank; pp_mailbox_commit then raises the flags
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
}

// loss + flags of one slot: after the records of the same slot (any stream, ordered before this launch)
__global__ void mailbox_commit_kernel(pp_mailbox mb, int64_t N, const float* __restrict__ loss) {
  const unsigned seq = mb.state[mb.slot] + 1u;
  const int64_t off = (static_cast<int64_t>(mb.slot) * mb.world + mb.rank) * mb.block_bytes;
  for (int p = threadIdx.x; p < mb.world; p += blockDim.x) {
    unsigned char* blk = static_cast<unsigned char*>(mb.peer_bufs[p]) + off;
    *reinterpret_cast<double*>(blk + loss_offset(N)) = loss ? static_cast<double>(*loss) : 0.0;
    st_release_sys(reinterpret_cast<unsigned*>(blk + loss_offset(N) + 8), seq);
  }
  __syncthreads();
  if (threadIdx.x == 0) mb.state[mb.slot] = seq;
}

__global__ void mailbox_wait_kernel(const unsigned char* __restrict__ local_buf, int world, int slot, int64_t block_bytes,
                                    int64_t flag_off, unsigned expected, long long timeout_cycles, int* __restrict__ status) {
  const int src = threadIdx.x;
  if (src >= world) return;
  const unsigned* flag = reinterpret_cast<const unsigned*>(local_buf + (static_cast<int64_t>(slot) * world + src) * block_bytes + flag_off);
  const long long t0 = clock64();
  // sequence numbers only grow; a later one means a newer step already overwrote the slot
  while (static_cast<int>(ld_acquire_sys(flag) - expected) < 0) {
    if (clock64() - t0 > timeout_cycles) {
      atomicExch(status, 1 + src);
      return;
    }
    __nanosleep(200);
  }
}

}  // namespace

extern "C" {

int64_t pp_mailbox_block_bytes(int64_t n_records) { return loss_offset(n_records) + 16; }

int pp_pack_records(int64_t N, const double* keypoints, const float* scores, const float* probabilities,
                    const float* visibilities, const float* oks, const float* errors, float inv_diagonal, double* records,
                    const pp_mailbox* mailbox, pp_stream_t stream) {
  PP_REQUIRE(N >= 0 && N < (1ll << 31), PP_ERR_INVALID_ARG, "pp_pack_records: bad N=%lld", static_cast<long long>(N));
  if (N == 0) return PP_OK;
  PP_REQUIRE(keypoints && scores && probabilities && visibilities && oks && errors, PP_ERR_INVALID_ARG,
             "pp_pack_records: null input");
  PP_REQUIRE(records || mailbox, PP_ERR_INVALID_ARG, "pp_pack_records: neither a local output nor a mailbox");
  pp_mailbox mb{};
  if (mailbox) {
    mb = *mailbox;
    PP_REQUIRE(mb.peer_bufs && mb.state && mb.world >= 1 && mb.rank >= 0 && mb.rank < mb.world && mb.slots >= 1 &&
                   mb.slot >= 0 && mb.slot < mb.slots && mb.block_bytes == pp_mailbox_block_bytes(N),
               PP_ERR_INVALID_ARG, "pp_pack_records: inconsistent mailbox (world=%d rank=%d slot=%d/%d block_bytes=%lld)",
               mb.world, mb.rank, mb.slot, mb.slots, static_cast<long long>(mb.block_bytes));
  }
  const int grid = static_cast<int>((N + kRecThreads - 1) / kRecThreads);
  pack_records_kernel<<<grid, kRecThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      N, keypoints, scores, probabilities, visibilities, oks, errors, inv_diagonal, records, mb);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

int pp_mailbox_commit(const pp_mailbox* mailbox, int64_t n_records, const float* loss, pp_stream_t stream) {
  PP_REQUIRE(mailbox != nullptr, PP_ERR_INVALID_ARG, "pp_mailbox_commit: null mailbox");
  const pp_mailbox mb = *mailbox;
  PP_REQUIRE(mb.peer_bufs && mb.state && mb.world >= 1 && mb.rank >= 0 && mb.rank < mb.world && mb.slots >= 1 && mb.slot >= 0 &&
                 mb.slot < mb.slots && mb.block_bytes == pp_mailbox_block_bytes(n_records),
             PP_ERR_INVALID_ARG, "pp_mailbox_commit: inconsistent mailbox");
  mailbox_commit_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(mb, n_records, loss);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

int pp_mailbox_wait(const void* local_mailbox, int32_t world, int32_t slot, int64_t n_records, uint32_t expected_seq,
                    int64_t timeout_us, int32_t* status, pp_stream_t stream) {
  PP_REQUIRE(local_mailbox && status && world >= 1 && world <= 1024 && slot >= 0 && n_records >= 0, PP_ERR_INVALID_ARG,
             "pp_mailbox_wait: bad argument");
  const long long cycles = static_cast<long long>(timeout_us) * 2000ll;   // ~2 GHz
  mailbox_wait_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const unsigned char*>(local_mailbox), world, slot, pp_mailbox_block_bytes(n_records), loss_offset(n_records) + 8,
      expected_seq, cycles, status);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

}  // extern "C"
