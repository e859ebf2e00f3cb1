// Pieces of the peer-memory push all-reduce (allreduce.cu) that the backward tail (render_stages.cu:
// post_backward_kernel's push role) shares with the stand-alone kernel.  Protocol: see allreduce.cu.
#pragma once
#include "trb_common.cuh"

namespace trb {

constexpr int kMaxSegments = 4;
constexpr int kMaxPeers = 16;
constexpr int kArChunk = 1024;  // elements per block

struct ArSegments {
  float* ptr[kMaxSegments];
  long long start[kMaxSegments + 1];  // prefix offsets; start[count] = total
  int count;
};

struct ArPeers {
  // rank r's inbox as mapped here: [2 parities][world senders][capacity] words of (epoch << 32 | value bits)
  unsigned long long* inbox[kMaxPeers];
};

// What the backward tail needs to push its finished gradients: the segments, the peers, and the counters.
struct ArPush {
  ArSegments seg;
  ArPeers peers;
  long long capacity;
  int rank, world;
  unsigned* epochs;     // per block of kArChunk elements (advanced by the receive kernel)
  int* error;
  unsigned* done;       // blocks of post_backward_kernel that have finished their share (reset by the receive kernel)
};

__device__ __forceinline__ void st_relaxed_sys_b64(unsigned long long* addr, unsigned long long w) {
  asm volatile("st.global.relaxed.sys.b64 [%0], %1;" ::"l"(addr), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_b64(const unsigned long long* addr) {
  unsigned long long w;
  asm volatile("ld.global.relaxed.sys.b64 %0, [%1];" : "=l"(w) : "l"(addr) : "memory");
  return w;
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* addr) {
  unsigned w;
  asm volatile("ld.global.acquire.gpu.u32 %0, [%1];" : "=r"(w) : "l"(addr) : "memory");
  return w;
}

// element i of the concatenated segments -> its address
__device__ __forceinline__ float* ar_element(const ArSegments& seg, long long i) {
  int s = 0;
#pragma unroll
  for (int k = 1; k < kMaxSegments; ++k) s += (k < seg.count && i >= seg.start[k]) ? 1 : 0;
  return seg.ptr[s] + (i - seg.start[s]);
}

// Push phase of block `blk` (256 threads): elements [blk * kArChunk, (blk + 1) * kArChunk) of the segments, tagged
// with the block's next epoch (epochs[blk] + 1), into every peer's inbox.  LDCG: the values may have been produced by other CTAs of
// the same kernel (L2 reductions).
__device__ __forceinline__ void ar_push_block(const ArSegments& seg, const ArPeers& p, int rank, int world,
                                              long long capacity, unsigned epoch, int blk) {
  constexpr int PER = kArChunk / 256;
  const long long total = seg.start[seg.count];
  const size_t parity_off = (size_t)(epoch & 1u) * world * capacity;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const long long i = (long long)blk * kArChunk + u * 256 + threadIdx.x;
    if (i < total) {
      const float v = __ldcg(ar_element(seg, i));
      const unsigned long long w = ((unsigned long long)epoch << 32) | __float_as_uint(v);
      for (int r = 0; r < world; ++r)
        if (r != rank) st_relaxed_sys_b64(p.inbox[r] + parity_off + (size_t)rank * capacity + i, w);
    }
  }
}

// Host side: fills the segment / peer tables from the C-ABI arguments; TRB_OK or TRB_ERR_BAD_ARG.
inline int ar_fill_tables(float* const* host_segments, const int64_t* host_counts, int num_segments,
                          void* const* host_peer_inbox, int64_t capacity_floats, int rank, int world,
                          ArSegments& seg, ArPeers& p, long long& total) {
  if (num_segments < 1 || num_segments > kMaxSegments || world < 1 || world > kMaxPeers || rank < 0 ||
      rank >= world || !host_segments || !host_counts || !host_peer_inbox || capacity_floats < 1)
    return TRB_ERR_BAD_ARG;
  seg.count = num_segments;
  total = 0;
  for (int i = 0; i < kMaxSegments; ++i) {
    seg.ptr[i] = i < num_segments ? host_segments[i] : nullptr;
    seg.start[i] = total;
    if (i < num_segments) {
      if (host_counts[i] < 0 || !host_segments[i]) return TRB_ERR_BAD_ARG;
      total += host_counts[i];
    }
  }
  for (int i = num_segments; i <= kMaxSegments; ++i) seg.start[i] = total;
  if (total > capacity_floats) return TRB_ERR_BAD_ARG;
  for (int r = 0; r < kMaxPeers; ++r) {
    p.inbox[r] = r < world ? (unsigned long long*)host_peer_inbox[r] : nullptr;
    if (r < world && !p.inbox[r]) return TRB_ERR_BAD_ARG;
  }
  return TRB_OK;
}

#ifdef TRB_STEP_STAMPS
// Diagnostic builds (TRB_EXTRA_NVCC_FLAGS=-DTRB_STEP_STAMPS): %globaltimer stamps of a step's kernels in a ring of
// 256 rows x 8 words -- [0] prep_kernel block 0 start, [1] first / [2] last post_backward_kernel block start / exit,
// [3] first all-reduce block past its griddepcontrol.wait / [4] last all-reduce block exit.  prep_kernel opens a row
// per step.  Every translation unit keeps its own copy of the two pointers (trb_debug_set_step_stamps sets both).
struct StepStamps { unsigned long long* ring; unsigned* step; };
__device__ __forceinline__ unsigned long long stamp_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long* stamp_row(const StepStamps& s) {
  return s.ring + (size_t)(*(volatile unsigned*)s.step & 255u) * 8;
}
int set_step_stamps_stages(unsigned long long* ring, unsigned* step);     // render_stages.cu
int set_step_stamps_allreduce(unsigned long long* ring, unsigned* step);  // allreduce.cu
#endif

// Launches the receive-and-sum half on its own (allreduce.cu); the push half ran inside post_backward_kernel.
int launch_allreduce_receive(const ArPush& a, cudaStream_t st);

}  // namespace trb
