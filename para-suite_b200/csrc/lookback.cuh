// Single-pass prefix over tiles ("decoupled look-back") for the pileup kernels.
//
// Every tile publishes its own aggregate as soon as it has it (kind 1) and, once it knows the prefix of all
// earlier tiles, the inclusive prefix (kind 2).  A tile finds its exclusive prefix by walking back over the
// descriptors of its predecessors, 32 at a time with one warp, combining aggregates until it meets an inclusive
// prefix.  The operator may be non-commutative (segmented reductions): operands are always combined in tile order.
// Tile numbers are handed out by an atomic counter when a block starts, so every predecessor of a running tile is
// itself running or finished and the walk cannot dead-lock.
// Payloads wider than one word are written with plain stores, then a fence, then the status word (release); readers
// load the status (acquire) before the payload.  The status carries the launch epoch so descriptors never need a reset.
#pragma once
#include <cstdint>

template <typename T>
struct alignas(16) LbDesc {
  T agg;
  T incl;
  unsigned int status;   // (epoch << 2) | kind
  unsigned int pad[3];
};

__device__ __forceinline__ unsigned int lb_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void lb_st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename T>
__device__ __forceinline__ T lb_shfl_down(const T& v, unsigned int d) {
  static_assert(sizeof(T) % 4 == 0, "payload must be a multiple of 4 bytes");
  T o;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* t = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (unsigned int k = 0; k < sizeof(T) / 4; ++k) t[k] = __shfl_down_sync(0xFFFFFFFFu, s[k], d);
  return o;
}
template <typename T>
__device__ __forceinline__ T lb_shfl_up(const T& v, unsigned int d) {
  T o;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* t = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (unsigned int k = 0; k < sizeof(T) / 4; ++k) t[k] = __shfl_up_sync(0xFFFFFFFFu, s[k], d);
  return o;
}
template <typename T>
__device__ __forceinline__ T lb_shfl(const T& v, int src) {
  T o;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* t = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (unsigned int k = 0; k < sizeof(T) / 4; ++k) t[k] = __shfl_sync(0xFFFFFFFFu, s[k], src);
  return o;
}

// payload loads go to L2 (the line may sit in this SM's L1 from an earlier poll of a neighbouring descriptor)
template <typename T>
__device__ __forceinline__ T lb_load(const T* p) {
  T o;
  const unsigned int* s = reinterpret_cast<const unsigned int*>(p);
  uint32_t* t = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (unsigned int k = 0; k < sizeof(T) / 4; ++k) t[k] = __ldcg(s + k);
  return o;
}

// one thread publishes
template <typename T>
__device__ __forceinline__ void lb_publish(LbDesc<T>* d, const T& v, unsigned int kind, unsigned int epoch) {
  if (kind == 1u) d->agg = v; else d->incl = v;
  __threadfence();
  lb_st_release(&d->status, (epoch << 2) | kind);
}

// Called by one full warp.  Returns (in every lane) op-combination of all tiles < tile, in tile order.
template <typename T, typename Op>
__device__ __forceinline__ T lb_exclusive_prefix(LbDesc<T>* descs, int tile, unsigned int epoch, Op op, const T& identity) {
  const int lane = threadIdx.x & 31;
  T acc = identity;
  int base = tile - 1;
  while (base >= 0) {
    const int p = base - lane;
    unsigned int kind = 2u;          // lanes before tile 0 behave like an inclusive prefix holding the identity
    T v = identity;
    if (p >= 0) {
      unsigned int st;
      do { st = lb_ld_acquire(&descs[p].status); } while ((st >> 2) != epoch || (st & 3u) == 0u);
      kind = st & 3u;
      v = lb_load(kind == 2u ? &descs[p].incl : &descs[p].agg);
    }
    const unsigned int incl_mask = __ballot_sync(0xFFFFFFFFu, kind == 2u);
    const int first = __ffs((int)incl_mask) - 1;      // nearest predecessor that already knows its inclusive prefix
    const int last = first >= 0 ? first : 31;
    if (lane > last) v = identity;
    // ordered reduction: lane 0 ends with v[last] op ... op v[1] op v[0]  (higher lane = earlier tile = left operand)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const T o = lb_shfl_down(v, d);
      if (lane + d < 32) v = op(o, v);
    }
    const T w = lb_shfl(v, 0);
    acc = op(w, acc);
    if (first >= 0) break;
    base -= 32;
  }
  return acc;
}
