// Device-side pieces of the peer-memory mailbox shared by pp_records.cu (record packing, commit, consumer side) and
// pp_loss.cu (the loss' finalize kernel publishes the step's loss itself).  Protocol: see pp_records.cu.
#pragma once

#include <cstdint>

#include "../../include/probpose_b200.h"

namespace pp_mailbox_dev {

__host__ __device__ inline int64_t loss_offset(int64_t n_records) { return (n_records * 56 + 15) / 16 * 16; }

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned* ack_word(void* base, const pp_mailbox& mb, int slot, int consumer) {
  return reinterpret_cast<unsigned*>(static_cast<unsigned char*>(base) + static_cast<int64_t>(mb.slots) * mb.world * mb.block_bytes) +
         slot * mb.world + consumer;
}

// One party of a slot's publication is done (its stores are fenced): the last of the two raises the flags.  Called by
// ALL 32 lanes of one warp: the flags of the `world` peers go out lane-parallel -- a release store waits for every earlier
// store of its thread to be acknowledged over NVLink, so one thread storing world flags in a row pays world round trips
// (measured: ~20 us of a 100 us step at world = 8), a warp pays one.
__device__ __forceinline__ void mailbox_arrive(const pp_mailbox& mb, int64_t N) {
  const int lane = threadIdx.x & 31;
  unsigned* seq_w = mb.state + mb.slot;
  unsigned* arrive_w = mb.state + mb.slots + mb.slot;
  unsigned first = 0;
  if (lane == 0) first = atomicAdd(arrive_w, 1u);
  first = __shfl_sync(0xffffffffu, first, 0);
  if (first != 1u) return;                      // the other party is still at work: it will publish
  if (lane == 0) {
    *arrive_w = 0u;
    __threadfence_system();                     // acquire side of the counter; the lanes' release stores below are cumulative
  }
  __syncwarp();
  const unsigned seq = *seq_w + 1u;
  const int64_t off = (static_cast<int64_t>(mb.slot) * mb.world + mb.rank) * mb.block_bytes;
  for (int p = lane; p < mb.world; p += 32)
    st_release_sys(reinterpret_cast<unsigned*>(static_cast<unsigned char*>(mb.peer_bufs[p]) + off + loss_offset(N) + 8), seq);
  __syncwarp();
  if (lane == 0) *seq_w = seq;
}

// the loss party (all 32 lanes of one warp, the same `loss` in every lane): store the step's local loss into the block on
// every rank, then arrive
__device__ __forceinline__ void mailbox_store_loss_and_arrive(const pp_mailbox& mb, int64_t N, double loss) {
  const int lane = threadIdx.x & 31;
  const int64_t off = (static_cast<int64_t>(mb.slot) * mb.world + mb.rank) * mb.block_bytes;
  for (int p = lane; p < mb.world; p += 32)
    *reinterpret_cast<double*>(static_cast<unsigned char*>(mb.peer_bufs[p]) + off + loss_offset(N)) = loss;
  __threadfence_system();
  __syncwarp();
  mailbox_arrive(mb, N);
}

}  // namespace pp_mailbox_dev
