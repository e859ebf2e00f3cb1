// One-shot sum all-reduce of the view-shared gradients over NVLink / NVSwitch peer memory.
//
// The path shards over views; the only exchange of a step is the sum of the per-rank partial gradients of the
// parameters all views share (vertices, vertex colours / texture map): 70 KB for the cow -- pure latency
// regime, where the NCCL route (cat + all-reduce + copy back) costs ~16 us of GPU time in a 250 us step.
//
// Push protocol, one kernel, ONE one-way NVLink latency: every rank writes each of its values, paired with
// the call's epoch in the same 8-byte word, straight into an inbox in every peer's memory (torch symmetric
// memory: each rank has all peers' inboxes mapped); the receiver polls its own inbox word by word until the
// epoch matches -- value and epoch travel as ONE 64-bit scalar (st/ld.relaxed.sys.b64: single-copy atomic in
// the PTX memory model, unlike a .v2.b32 pair), so the flag validates the value it arrives with and no fence
// or barrier is needed -- and sums the contributions in rank order (every rank gets the
// bit-identical result).  Inboxes are double-buffered by epoch parity: a sender rewrites a parity only two
// calls later, after it has received the intervening call's data from that peer, i.e. after the peer finished
// reading.  Block b always owns elements [b*1024, (b+1)*1024) and keeps its own epoch counter in device
// memory, so the kernel can be replayed from a CUDA graph and the segment layout may change between calls.
#include "trb_internal.cuh"
#include "stages.cuh"   // pdl_wait
#include "allreduce.cuh"

namespace trb {

// Optional device-side timing of the protocol (trb_allreduce_set_timing): block 0 accumulates, in nanoseconds of
// %globaltimer, [0] push time, [1] wait-and-sum time (rank skew + one NVLink one-way latency), [2] call count.
static unsigned long long* g_ar_timing = nullptr;

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

#ifdef TRB_STEP_STAMPS
static __device__ StepStamps g_stamps_ar = {nullptr, nullptr};
int set_step_stamps_allreduce(unsigned long long* ring, unsigned* step) {
  StepStamps s = {ring, step};
  return cudaMemcpyToSymbol(g_stamps_ar, &s, sizeof(s)) == cudaSuccess ? TRB_OK : TRB_ERR_CUDA;
}
#endif

__global__ void __launch_bounds__(256)
allreduce_push_kernel(const ArSegments seg, const ArPeers p, int rank, int world, long long capacity,
                      unsigned* epochs, int* error, unsigned long long* timing) {
  constexpr int PER = kArChunk / 256;
  // launched as a programmatic dependent launch: resident while the producer of the gradients drains
  pdl_wait();
#ifdef TRB_STEP_STAMPS
  unsigned long long* stamp = g_stamps_ar.ring ? stamp_row(g_stamps_ar) : nullptr;
  if (stamp && threadIdx.x == 0) atomicMin(stamp + 3, stamp_now());
#endif
  const bool timed = timing != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  const unsigned long long t0 = timed ? global_ns() : 0ull;
  const unsigned epoch = epochs[blockIdx.x] + 1u;
  const long long total = seg.start[seg.count];
  const size_t parity_off = (size_t)(epoch & 1u) * world * capacity;
  float v[PER];
  float* dst[PER];
  // ---- 1. push my values (+ epoch) into every peer's inbox
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const long long i = (long long)blockIdx.x * kArChunk + u * 256 + threadIdx.x;
    dst[u] = nullptr; v[u] = 0.0f;
    if (i < total) {
      dst[u] = ar_element(seg, i);
      v[u] = *dst[u];
      for (int r = 0; r < world; ++r)
        if (r != rank)
          st_relaxed_sys_b64(p.inbox[r] + parity_off + (size_t)rank * capacity + i,
                             ((unsigned long long)epoch << 32) | __float_as_uint(v[u]));
    }
  }
  const unsigned long long t1 = timed ? global_ns() : 0ull;
  // ---- 2. collect the peers' values from my inbox, sum in rank order.  All peers' words of an element are
  // requested BEFORE the first one is looked at: the system-scope loads are then in flight together (polling the
  // peers one after the other cost ~1.5 us each -- 11 us of wait at 8 GPUs with every contribution already there).
  const unsigned long long* mine = p.inbox[rank] + parity_off;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    if (dst[u] == nullptr) continue;
    const long long i = (long long)blockIdx.x * kArChunk + u * 256 + threadIdx.x;
    float acc = 0.0f;
    bool ok = true;
    for (int r0 = 0; r0 < world && ok; r0 += 8) {   // peers in groups of 8 (one group on an 8-GPU box)
      unsigned long long w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = r0 + j;
        if (r < world && r != rank) w[j] = ld_relaxed_sys_b64(mine + (size_t)r * capacity + i);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = r0 + j;
        if (r >= world || !ok) continue;
        if (r == rank) { acc += v[u]; continue; }
        const unsigned long long* src = mine + (size_t)r * capacity + i;
        unsigned spins = 0;
        while ((unsigned)(w[j] >> 32) != epoch) {
          if (++spins > (1u << 26)) { atomicExch(error, 1); ok = false; break; }
          w[j] = ld_relaxed_sys_b64(src);
        }
        acc += __uint_as_float((unsigned)w[j]);
      }
    }
    // a peer that never arrived: the partial sum is NOT written (the local gradient stays as it was and the
    // error flag tells the host -- PeerAllReduce.check())
    if (ok) *dst[u] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) epochs[blockIdx.x] = epoch;
  if (timed) {
    const unsigned long long t2 = global_ns();
    timing[0] += t1 - t0; timing[1] += t2 - t1; timing[2] += 1ull;
  }
#ifdef TRB_STEP_STAMPS
  if (stamp && threadIdx.x == 0) atomicMax(stamp + 4, stamp_now());
#endif
}

// The receive-and-sum half on its own: the push half ran inside post_backward_kernel (render_stages.cu), whose last
// blocks wrote this rank's finished gradients into the peers' inboxes while the rest of that kernel drained.  Same
// inbox, epochs and rank-ordered sum as allreduce_push_kernel, so fused and stand-alone calls can alternate.
__global__ void __launch_bounds__(256)
allreduce_receive_kernel(const ArSegments seg, const ArPeers p, int rank, int world, long long capacity,
                         unsigned* epochs, int* error, unsigned* done, unsigned long long* timing) {
  constexpr int PER = kArChunk / 256;
  pdl_wait();
  const bool timed = timing != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  const unsigned long long t1 = timed ? global_ns() : 0ull;
  const unsigned epoch = epochs[blockIdx.x] + 1u;
  const long long total = seg.start[seg.count];
  const size_t parity_off = (size_t)(epoch & 1u) * world * capacity;
  const unsigned long long* mine = p.inbox[rank] + parity_off;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const long long i = (long long)blockIdx.x * kArChunk + u * 256 + threadIdx.x;
    if (i >= total) continue;
    float* dst = ar_element(seg, i);
    const float own = *dst;
    float acc = 0.0f;
    bool ok = true;
    for (int r0 = 0; r0 < world && ok; r0 += 8) {
      unsigned long long w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = r0 + j;
        if (r < world && r != rank) w[j] = ld_relaxed_sys_b64(mine + (size_t)r * capacity + i);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = r0 + j;
        if (r >= world || !ok) continue;
        if (r == rank) { acc += own; continue; }
        const unsigned long long* src = mine + (size_t)r * capacity + i;
        unsigned spins = 0;
        while ((unsigned)(w[j] >> 32) != epoch) {
          if (++spins > (1u << 26)) { atomicExch(error, 1); ok = false; break; }
          w[j] = ld_relaxed_sys_b64(src);
        }
        acc += __uint_as_float((unsigned)w[j]);
      }
    }
    if (ok) *dst = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    epochs[blockIdx.x] = epoch;
    if (blockIdx.x == 0) *done = 0u;   // the next post_backward_kernel counts its blocks from zero again
  }
  if (timed) {
    const unsigned long long t2 = global_ns();
    timing[1] += t2 - t1; timing[2] += 1ull;
  }
}

int launch_allreduce_receive(const ArPush& a, cudaStream_t st) {
  const long long total = a.seg.start[a.seg.count];
  if (total == 0) return TRB_OK;
  const int grid = (int)((total + kArChunk - 1) / kArChunk);
  TRB_CUDA_TRY(launch_pdl(allreduce_receive_kernel, dim3(grid), dim3(256), 0, st, a.seg, a.peers, a.rank, a.world,
                          a.capacity, a.epochs, a.error, a.done, g_ar_timing));
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

}  // namespace trb

using namespace trb;

/* number of per-block epoch counters a staging area of `capacity_floats` needs */
extern "C" int trb_allreduce_grid(int64_t capacity_floats) {
  return (int)((capacity_floats + kArChunk - 1) / kArChunk);
}

extern "C" int trb_allreduce_sum_f32(float* const* host_segments, const int64_t* host_counts, int num_segments,
                                     void* const* host_peer_inbox, int64_t capacity_floats, int rank, int world,
                                     uint32_t* epochs, int32_t* error_flag, int device, trb_stream_t stream) {
  if (!error_flag || !epochs) return TRB_ERR_BAD_ARG;
  ArSegments seg;
  ArPeers p;
  long long total = 0;
  const int rc = ar_fill_tables(host_segments, host_counts, num_segments, host_peer_inbox, capacity_floats, rank, world,
                                seg, p, total);
  if (rc != TRB_OK) return rc;
  if (total == 0) return TRB_OK;
  TRB_ENTER(device);
  const int grid = (int)((total + kArChunk - 1) / kArChunk);
  TRB_CUDA_TRY(launch_pdl(allreduce_push_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, seg, p, rank, world,
                          (long long)capacity_floats, (unsigned*)epochs, (int*)error_flag, g_ar_timing));
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

#ifdef TRB_STEP_STAMPS
/* diagnostic builds: ring u64[256][8] + step counter u32[1] (device, zeroed by the caller); NULLs switch it off */
extern "C" int trb_debug_set_step_stamps(uint64_t* ring, uint32_t* step) {
  int rc = set_step_stamps_stages((unsigned long long*)ring, (unsigned*)step);
  if (rc != TRB_OK) return rc;
  return set_step_stamps_allreduce((unsigned long long*)ring, (unsigned*)step);
}
#endif

/* diagnostics: `device_u64x3` (zeroed by the caller) accumulates block 0's push / wait-and-sum nanoseconds and the
 * call count of every later trb_allreduce_sum_f32 launch; NULL switches it off.  Process-wide. */
extern "C" int trb_allreduce_set_timing(uint64_t* device_u64x3) {
  g_ar_timing = (unsigned long long*)device_u64x3;
  return TRB_OK;
}
