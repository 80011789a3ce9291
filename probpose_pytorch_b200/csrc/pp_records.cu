// Decoded-keypoint records and their exchange between GPUs.
//
//  * pp_pack_records : the tail of Codec.decode (codec.py:249-263) on the device -- one packed (N, 7) float64 record
//                      (x, y, score, probability, visibility, oks, error / diagonal) per keypoint, the unit that is
//                      gathered across GPUs -- written in ONE kernel, and, when a mailbox is given, to
//                      every peer GPU in the same kernel: the final keypoint gather is plain stores into peer memory
//                      over NVLink (each rank writes its block into every rank's mailbox).  No collective call, no
//                      intermediate copy; latency-bound (a few hundred KB per rank).
//  * pp_mailbox_commit: stores the step's local loss next to the records (pp_oks_loss_forward[_encoded] does the same
//                      from its own finalize kernel when it is given the mailbox).
//  * publication     : a slot is published -- its per-source flag raised on every rank with a system-scope release
//                      store -- by whichever of the two parties (record packing, loss) finishes LAST: both arrive on
//                      a device-side counter, so no launch has to be ordered after both of them and the step's two
//                      streams need no join for the exchange.
//  * pp_mailbox_wait / pp_mailbox_ack : the consumer side -- wait (bounded) until every source rank's flag of a slot
//                      has reached the expected sequence number; after using the blocks, acknowledge the sequence
//                      number to every producer.
//  * flow control    : with pp_mailbox.flow_control a producer does not touch a slot before every consumer has
//                      acknowledged the slot's previous publication (bounded wait in pp_pack_records), so a reader can
//                      never see a block change under it: no torn records, no mixing of steps.  Every rank must then
//                      consume (wait + ack) every publication.
//
// Mailbox layout (identical on every rank, symmetric allocation): slots x world blocks of block_bytes
//   [ N * 7 doubles | pad to 16 | loss (double) | flag (uint32) | pad ]   with block_bytes = round_up(N * 56, 16) + 16,
// followed by slots x world uint32 acknowledgements (ack[slot][consumer], written by the consumers over NVLink).
// Local state (pp_mailbox.state, 5 * slots + 1 uint32, zero before the first use): sequence number, arrival counter,
// finished-block counter of the packing kernel, consumed sequence number and finished-block counter of the consumer
// kernel of every slot, then one status word (1 + rank of a consumer that did not acknowledge in time).  The consumed
// counters make the consumer side (pp_mailbox_consume) free of host-provided sequence numbers, so it can be captured
// into the same CUDA graph as the step.
#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/probpose_b200.h"
#include "pp_common.cuh"
#include "pp_mailbox.cuh"

namespace {

using namespace pp_mailbox_dev;

constexpr int kRecThreads = 256;

__global__ void __launch_bounds__(kRecThreads)
pack_records_kernel(int64_t N, const double* __restrict__ keypoints, const float* __restrict__ scores,
                    const float* __restrict__ prob, const float* __restrict__ vis, const float* __restrict__ oks,
                    const float* __restrict__ err, float inv_diag, double* __restrict__ rec, pp_mailbox mb,
                    long long ack_timeout_cycles) {
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
  if (mb.peer_bufs && mb.flow_control && threadIdx.x < mb.world) {
    // flow control: consumer `threadIdx.x` must have acknowledged this slot's previous publication (sequence number
    // state[slot]) before its copy of the block is overwritten.  Normally long true: the slot was published `slots`
    // steps ago.
    const unsigned want = mb.state[mb.slot];
    const unsigned* ack = ack_word(mb.peer_bufs[mb.rank], mb, mb.slot, threadIdx.x);
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys(ack) - want) < 0) {
      if (clock64() - t0 > ack_timeout_cycles) {
        atomicCAS(mb.state + 5 * mb.slots, 0u, 1u + threadIdx.x);   // reported by the next status check; go on
        break;
      }
      __nanosleep(100);
    }
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
  // one system-scope fence per block, after the block barrier (cumulative over the stores the barrier has ordered); the
  // block that finishes last is this kernel's arrival at the slot's publication
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned last = 0;
    if (threadIdx.x == 0) {
      __threadfence_system();
      unsigned* done_w = mb.state + 2 * mb.slots + mb.slot;
      last = atomicAdd(done_w, 1u) == gridDim.x - 1;
      if (last) *done_w = 0u;
    }
    if (__shfl_sync(0xffffffffu, last, 0)) mailbox_arrive(mb, N);   // warp-wide: the flags go out lane-parallel
  }
}

__global__ void mailbox_commit_kernel(pp_mailbox mb, int64_t N, const float* __restrict__ loss) {   // <<<1, 32>>>
  mailbox_store_loss_and_arrive(mb, N, loss ? static_cast<double>(*loss) : 0.0);
}

// the loss party, one step late: only if the records party of the slot has arrived and is waiting (nothing otherwise)
__global__ void mailbox_commit_deferred_kernel(pp_mailbox mb, int64_t N, const float* __restrict__ loss) {   // <<<1, 32>>>
  unsigned arrived = 0;
  if (threadIdx.x == 0) arrived = *reinterpret_cast<volatile unsigned*>(mb.state + mb.slots + mb.slot);
  if (__shfl_sync(0xffffffffu, arrived, 0) != 1u) return;
  mailbox_store_loss_and_arrive(mb, N, loss ? static_cast<double>(*loss) : 0.0);
}

__global__ void mailbox_ack_kernel(pp_mailbox mb, unsigned seq) {
  const int p = threadIdx.x;
  if (p < mb.world) st_release_sys(ack_word(mb.peer_bufs[p], mb, mb.slot, mb.rank), seq);
}

__global__ void mailbox_wait_kernel(const unsigned char* __restrict__ local_buf, int world, int slot, int64_t block_bytes,
                                    int64_t flag_off, unsigned expected, long long timeout_cycles, int* __restrict__ status) {
  const int src = threadIdx.x;
  if (src >= world) return;
  const unsigned* flag = reinterpret_cast<const unsigned*>(local_buf + (static_cast<int64_t>(slot) * world + src) * block_bytes + flag_off);
  const long long t0 = clock64();
  unsigned seen;
  while (static_cast<int>((seen = ld_acquire_sys(flag)) - expected) < 0) {
    if (clock64() - t0 > timeout_cycles) {
      atomicCAS(status, 0, 1 + src);
      return;
    }
    __nanosleep(200);
  }
  // sequence numbers only grow: a later one means the producer has already overwritten the block (possible only
  // without flow control) -- an error, not a success
  if (seen != expected) atomicCAS(status, 0, -(1 + src));
}

// Consumer side with device-side sequence numbers, one kernel: the next publication of mb.slot that this rank has not
// consumed yet (state[3 S + slot] + 1).  Nothing to do when this rank itself has not published it yet (every rank
// publishes every slot equally often, so "nothing new here" means "nothing new anywhere" at this point of the step
// sequence).  Every block waits for the flags of all sources itself (local loads), copies its share of the slot into the
// compact private buffers, and the block that finishes last acknowledges to every producer and advances the counter --
// the other blocks have read the counter by then, they are done.
constexpr int kConsumeThreads = 256;

template <typename V>
__device__ __forceinline__ void copy_unrolled(V* __restrict__ dst, const V* __restrict__ src, int64_t n, int64_t first,
                                              int64_t stride) {
  for (int64_t i = first; i < n; i += 8 * stride) {
    V v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i + u * stride < n) v[u] = src[i + u * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i + u * stride < n) dst[i + u * stride] = v[u];
  }
}

__global__ void __launch_bounds__(kConsumeThreads)
mailbox_consume_kernel(pp_mailbox mb, int64_t N, double* __restrict__ rec_out, double* __restrict__ loss_out,
                       long long timeout_cycles, int* __restrict__ status) {
  unsigned* consumed = mb.state + 3 * mb.slots + mb.slot;
  const unsigned expected = *consumed + 1u;
  if (static_cast<int>(mb.state[mb.slot] - expected) < 0) return;
  const unsigned char* slot_base =
      static_cast<const unsigned char*>(mb.peer_bufs[mb.rank]) + static_cast<int64_t>(mb.slot) * mb.world * mb.block_bytes;
  const int64_t flag_off = loss_offset(N) + 8;
  for (int src = threadIdx.x; src < mb.world; src += kConsumeThreads) {
    const unsigned* flag = reinterpret_cast<const unsigned*>(slot_base + src * mb.block_bytes + flag_off);
    const long long t0 = clock64();
    unsigned seen;
    bool ok = true;
    while (static_cast<int>((seen = ld_acquire_sys(flag)) - expected) < 0) {
      if (clock64() - t0 > timeout_cycles) {
        atomicCAS(status, 0, 1 + src);
        ok = false;
        break;
      }
      __nanosleep(200);
    }
    if (ok && seen != expected) atomicCAS(status, 0, -(1 + src));   // overwritten already (no flow control)
  }
  __syncthreads();
  if (rec_out) {
    // eight independent loads in flight per thread: the blocks were written by the peers over NVLink, every load is
    // an HBM round trip
    const int64_t per_src = N * 7;                    // doubles
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kConsumeThreads;
    const int64_t first = static_cast<int64_t>(blockIdx.x) * kConsumeThreads + threadIdx.x;
    if ((per_src & 1) == 0) {
      for (int src = 0; src < mb.world; ++src)
        copy_unrolled(reinterpret_cast<double2*>(rec_out + src * per_src),
                      reinterpret_cast<const double2*>(slot_base + src * mb.block_bytes), per_src >> 1, first, stride);
    } else {
      for (int src = 0; src < mb.world; ++src)
        copy_unrolled(rec_out + src * per_src, reinterpret_cast<const double*>(slot_base + src * mb.block_bytes), per_src,
                      first, stride);
    }
  }
  if (loss_out && blockIdx.x == 0)
    for (int src = threadIdx.x; src < mb.world; src += kConsumeThreads)
      loss_out[src] = *reinterpret_cast<const double*>(slot_base + src * mb.block_bytes + loss_offset(N));
  __shared__ int is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned* done_w = mb.state + 4 * mb.slots + mb.slot;
    is_last = atomicAdd(done_w, 1u) == gridDim.x - 1;
    if (is_last) {
      *done_w = 0u;
      __threadfence();
    }
  }
  __syncthreads();
  if (!is_last) return;
  if (mb.flow_control)     // one acknowledgement per producer, thread-parallel (a release store = one NVLink round trip)
    for (int p = threadIdx.x; p < mb.world; p += kConsumeThreads)
      st_release_sys(ack_word(mb.peer_bufs[p], mb, mb.slot, mb.rank), expected);
  if (threadIdx.x == 0) *consumed = expected;
}

}  // namespace

extern "C" {

static int check_consumer_mailbox(const char* fn, const pp_mailbox* mailbox, int64_t n_records);

int64_t pp_mailbox_block_bytes(int64_t n_records) { return loss_offset(n_records) + 16; }

int64_t pp_mailbox_bytes(int64_t n_records, int32_t world, int32_t slots) {
  if (n_records < 0 || world < 1 || slots < 1) return 0;
  return static_cast<int64_t>(slots) * world * pp_mailbox_block_bytes(n_records) + static_cast<int64_t>(slots) * world * 4;
}

int64_t pp_mailbox_state_words(int32_t slots) { return slots < 1 ? 0 : 5 * static_cast<int64_t>(slots) + 1; }

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
                   mb.slot >= 0 && mb.slot < mb.slots && mb.block_bytes == pp_mailbox_block_bytes(N) && mb.world <= kRecThreads,
               PP_ERR_INVALID_ARG, "pp_pack_records: inconsistent mailbox (world=%d rank=%d slot=%d/%d block_bytes=%lld)",
               mb.world, mb.rank, mb.slot, mb.slots, static_cast<long long>(mb.block_bytes));
  }
  const int grid = static_cast<int>((N + kRecThreads - 1) / kRecThreads);
  const long long ack_cycles = static_cast<long long>(pp_env_int("PP_MAILBOX_ACK_TIMEOUT_US", 2000000)) * 2000ll;   // ~2 GHz
  pack_records_kernel<<<grid, kRecThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      N, keypoints, scores, probabilities, visibilities, oks, errors, inv_diagonal, records, mb, ack_cycles);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

int pp_mailbox_commit(const pp_mailbox* mailbox, int64_t n_records, const float* loss, pp_stream_t stream) {
  PP_REQUIRE(mailbox != nullptr, PP_ERR_INVALID_ARG, "pp_mailbox_commit: null mailbox");
  const pp_mailbox mb = *mailbox;
  PP_REQUIRE(mb.peer_bufs && mb.state && mb.world >= 1 && mb.rank >= 0 && mb.rank < mb.world && mb.slots >= 1 && mb.slot >= 0 &&
                 mb.slot < mb.slots && mb.block_bytes == pp_mailbox_block_bytes(n_records),
             PP_ERR_INVALID_ARG, "pp_mailbox_commit: inconsistent mailbox");
  mailbox_commit_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(mb, n_records, loss);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

int pp_mailbox_commit_deferred(const pp_mailbox* mailbox, int64_t n_records, const float* loss, pp_stream_t stream) {
  if (int rc = check_consumer_mailbox("pp_mailbox_commit_deferred", mailbox, n_records)) return rc;
  mailbox_commit_deferred_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*mailbox, n_records, loss);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

int pp_mailbox_ack(const pp_mailbox* mailbox, uint32_t seq, pp_stream_t stream) {
  PP_REQUIRE(mailbox != nullptr, PP_ERR_INVALID_ARG, "pp_mailbox_ack: null mailbox");
  const pp_mailbox mb = *mailbox;
  PP_REQUIRE(mb.peer_bufs && mb.world >= 1 && mb.world <= 1024 && mb.rank >= 0 && mb.rank < mb.world && mb.slots >= 1 &&
                 mb.slot >= 0 && mb.slot < mb.slots,
             PP_ERR_INVALID_ARG, "pp_mailbox_ack: inconsistent mailbox");
  mailbox_ack_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(mb, seq);
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

static int check_consumer_mailbox(const char* fn, const pp_mailbox* mailbox, int64_t n_records) {
  PP_REQUIRE(mailbox != nullptr, PP_ERR_INVALID_ARG, "%s: null mailbox", fn);
  const pp_mailbox& mb = *mailbox;
  PP_REQUIRE(mb.peer_bufs && mb.state && mb.world >= 1 && mb.world <= 1024 && mb.rank >= 0 && mb.rank < mb.world && mb.slots >= 1 &&
                 mb.slot >= 0 && mb.slot < mb.slots && mb.block_bytes == pp_mailbox_block_bytes(n_records),
             PP_ERR_INVALID_ARG, "%s: inconsistent mailbox", fn);
  return PP_OK;
}

int pp_mailbox_consume(const pp_mailbox* mailbox, int64_t n_records, double* records_out, double* losses_out,
                       int64_t timeout_us, int32_t* status, pp_stream_t stream) {
  if (int rc = check_consumer_mailbox("pp_mailbox_consume", mailbox, n_records)) return rc;
  PP_REQUIRE(status != nullptr, PP_ERR_INVALID_ARG, "pp_mailbox_consume: null status");
  // a few blocks are enough: at most world x n_records x 56 bytes (tens of MB), beside the step's own kernels
  const int64_t doubles = records_out ? n_records * 7 * mailbox->world : 0;
  const int grid = static_cast<int>(std::min<int64_t>(64, std::max<int64_t>(1, doubles / (kConsumeThreads * 8))));
  mailbox_consume_kernel<<<grid, kConsumeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      *mailbox, n_records, records_out, losses_out, static_cast<long long>(timeout_us) * 2000ll, status);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

}  // extern "C"
