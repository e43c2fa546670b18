// Single-pass prefix over tiles ("decoupled look-back") for the pileup kernels.
//
// Every tile publishes its own aggregate as soon as it has it (kind 1) and, once it knows the prefix of all
// earlier tiles, its inclusive prefix (kind 2).  A tile finds its exclusive prefix by walking back over the
// descriptors of its predecessors, 32 at a time with one warp, combining aggregates until it meets an inclusive
// prefix.  Tile numbers are handed out by an atomic counter when a block starts, so every predecessor of a running
// tile is itself running or finished and the walk cannot dead-lock.
// A descriptor is ONE aligned 16-byte word {value, status}: it is written and read with single 128-bit accesses,
// so value and status always belong together and no fence is needed.  The status carries the launch epoch, so the
// descriptor array never needs a reset between launches.
#pragma once
#include <cstdint>

struct alignas(16) LbDesc {
  unsigned long long val;
  unsigned int status;   // (epoch << 2) | kind;  kind 0 = nothing yet; the epoch is 30 bits wide (next_epoch, pileup.cu)
  unsigned int pad;
};

__device__ __forceinline__ void lb_publish(LbDesc* d, unsigned long long v, unsigned int kind, unsigned int epoch) {
  const unsigned int st = (epoch << 2) | kind;
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(d), "r"((unsigned int)v),
               "r"((unsigned int)(v >> 32)), "r"(st), "r"(0u)
               : "memory");
}
__device__ __forceinline__ void lb_load(const LbDesc* d, unsigned long long& v, unsigned int& st) {
  unsigned int a, b, c, e;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(e) : "l"(d) : "memory");
  (void)e;
  v = ((unsigned long long)b << 32) | a;
  st = c;
}

struct LbMax {
  __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a > b ? a : b; }
};
struct LbSum {
  __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a + b; }
};

// Called by one full warp.  Returns (in every lane) the combination of all tiles < tile (operators are commutative).
template <typename Op>
__device__ __forceinline__ unsigned long long lb_exclusive_prefix(const LbDesc* descs, int tile, unsigned int epoch, Op op,
                                                                   unsigned long long identity) {
  const int lane = threadIdx.x & 31;
  unsigned long long acc = identity;
  int base = tile - 1;
  while (base >= 0) {
    const int p = base - lane;
    unsigned int kind = 2u;          // lanes before tile 0 behave like an inclusive prefix holding the identity
    unsigned long long v = identity;
    if (p >= 0) {
      unsigned int st;
      do { lb_load(&descs[p], v, st); } while ((st >> 2) != epoch || (st & 3u) == 0u);
      kind = st & 3u;
    }
    const unsigned int incl_mask = __ballot_sync(0xFFFFFFFFu, kind == 2u);
    const int first = __ffs((int)incl_mask) - 1;      // nearest predecessor that already knows its inclusive prefix
    const int last = first >= 0 ? first : 31;
    if (lane > last) v = identity;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
      v = op(v, o);
    }
    acc = op(acc, v);
    if (first >= 0) break;
    base -= 32;
  }
  return acc;
}
