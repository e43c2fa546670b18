// Fast path of the error-profile kernel (included by profile.cu inside its anonymous namespace).
//
// Shape: uniform read length L <= 64, one M/=/X cigar op per read (L == R) -- the PAR-CLIP case.
//
// Work unit = a WARP-tile of 64 consecutive reads (2 per lane).  Every warp runs its own 3-stage ring of
// cp.async.bulk (TMA) copies with its own mbarriers, so there is no block-wide barrier in the main loop:
// a warp that meets a read with many mismatches delays nobody else.  All per-base work is bit-parallel on
// 2-bit packed words:
//   match counts   per-thread bit-sliced ("vertical") counters over one-hot (A|C, G|T) x position lanes, merged
//                  across the warp with a carry-save butterfly every 2^P-2 reads (no atomics on the hot path)
//   mismatches     rare path: two native 32-bit shared atomics (count, quality) per mismatching base
//   quality sums   dp4a of the quality bytes against 0/-1 byte masks built with PRMT from the read codes,
//                  summed over ALL positions by read base and corrected by the mismatch / invalid sums at the end
// Reads outside the shape (flags, other cigars, contig edges) are appended to a dense list for
// profile_deferred_kernel instead of being walked inline (one slow lane would stall the other 31).

#define FAST_STAGES 2
#define WT_READS 64            // reads per warp-tile
#define FAST_THREADS 128
#define FAST_WARPS (FAST_THREADS / 32)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine), completion signalled on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ int dp4a_ss(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

struct FastStage {      // byte offsets inside one warp-stage buffer
  uint32_t meta, start, cigar, bases, qual, total;
};
__host__ __device__ inline FastStage fast_stage_layout(uint32_t L) {
  FastStage s;
  const uint32_t bpr = (L + 3) / 4;
  s.meta = 0;
  s.start = WT_READS * 4;
  s.cigar = 2 * WT_READS * 4;
  s.bases = 3 * WT_READS * 4;
  s.qual = s.bases + ((WT_READS * bpr + 15) & ~15u) + 16;   // +16: word reads may run past a row end
  s.total = (s.qual + WT_READS * L + 16 + 127) & ~127u;
  return s;
}

struct FastLayout {     // byte offsets of the block's shared memory
  uint32_t misc, mm_cnt, mm_q, fast, warp0, warp_stride, inv, bars, stage0, total;
};
__host__ __device__ inline FastLayout fast_layout(uint32_t max_len, uint32_t L, uint32_t nw) {
  FastLayout f;
  f.misc = 0;                                   // u64[16]
  f.mm_cnt = 128;                               // u32[max_len*16]
  f.mm_q = f.mm_cnt + max_len * 64;             // u32[max_len*16]
  f.fast = f.mm_q + max_len * 64;               // u32[max_len*4]
  f.warp0 = (f.fast + max_len * 16 + 127) & ~127u;
  f.bars = 0;                                   // per warp: u64[FAST_STAGES] (+pad to 32)
  f.inv = 32;                                   // per warp: u32[WT_READS*nw]
  f.stage0 = (32 + WT_READS * nw * 4 + 127) & ~127u;
  f.warp_stride = f.stage0 + FAST_STAGES * fast_stage_layout(L).total;
  f.total = f.warp0 + FAST_WARPS * f.warp_stride;
  return f;
}

// bit-sliced counter: planes[p] holds bit p of 32 independent counters
template <int NPL>
__device__ __forceinline__ void vc_add2(uint32_t (&pl)[NPL], uint32_t x, uint32_t y) {
  uint32_t s = pl[0] ^ x ^ y;
  uint32_t c = (pl[0] & x) | (pl[0] & y) | (x & y);
  pl[0] = s;
#pragma unroll
  for (int p = 1; p < NPL; ++p) {
    const uint32_t t = pl[p] & c;
    pl[p] ^= c;
    c = t;
  }
}

template <int NPL>
struct Planes { uint32_t p[NPL]; };

// sum the bit-sliced counters of the 32 lanes; lane l ends up with the integer total of bit-lane l
// (passed by value so the caller's planes stay in registers across the call)
template <int NPL>
__device__ __noinline__ uint32_t vc_warp_total(Planes<NPL> pl) {
  uint32_t a[NPL + 5];
#pragma unroll
  for (int p = 0; p < NPL; ++p) a[p] = pl.p[p];
#pragma unroll
  for (int p = NPL; p < NPL + 5; ++p) a[p] = 0;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int np = NPL + s;   // planes holding data before this step
    uint32_t c = 0;
#pragma unroll
    for (int p = 0; p < NPL + 5; ++p) {
      if (p <= np) {
        const uint32_t b = p < np ? __shfl_xor_sync(0xFFFFFFFFu, a[p], 1 << s) : 0u;
        const uint32_t av = a[p];
        a[p] = av ^ b ^ c;
        c = (av & b) | (av & c) | (b & c);
      }
    }
  }
  const uint32_t lane = threadIdx.x & 31;
  uint32_t tot = 0;
#pragma unroll
  for (int p = 0; p < NPL + 5; ++p) tot |= ((a[p] >> lane) & 1u) << p;
  return tot;
}

struct FastSmem {
  unsigned long long* s_misc;   // [0..3] sum of quality by read base over all positions, [4..7] same at invalid positions, [8] fast reads
  uint32_t* s_mm_cnt;           // [max_len*16] mismatch counts
  uint32_t* s_mm_q;             // [max_len*16] mismatch quality sums (two's complement)
  uint32_t* s_fast;             // [max_len*4] match counts by (position, base)
};

// match count of base a at one position from the four counted lanes E0..E3 (see fast_read)
__device__ __forceinline__ uint32_t fast_base_count(const uint32_t* e, uint32_t a) {
  const uint32_t e0 = e[0], e1 = e[1], e2 = e[2], e3 = e[3];
  return a == 0 ? e0 - e1 - e2 + e3 : (a == 1 ? e1 - e3 : (a == 2 ? e2 - e3 : e3));
}

// number of bit-sliced counter words for NW 2-bit words: (A|C) and (G|T) per word; when the last word holds at
// most 8 positions its two one-hot words share one counter word (A|C in the low half, G|T in the high half)
__host__ __device__ constexpr int fast_nc(int NW, int LT) {
  return (LT > 0 && LT - 16 * (NW - 1) <= 8) ? 2 * NW - 1 : 2 * NW;
}

// One read of the fast shape.  Produces the one-hot match words cw[] for the caller's bit-sliced counters.
// LT > 0: the (uniform) read length is a compile-time constant.
template <int NW, int LT>
__device__ __forceinline__ void fast_read(const DeviceRef& ref, const FastSmem& F, uint32_t Lrt, uint32_t g0, bool rev,
                                          const uint32_t* __restrict__ brow_w, uint32_t bshift,
                                          const uint32_t* __restrict__ qrow_w, uint32_t qshift,
                                          const unsigned char* __restrict__ qrow_b, const uint32_t (&lenmask)[NW],
                                          uint32_t* __restrict__ inv_row, bool has_n, const uint32_t (&tbl)[4],
                                          uint32_t (&cw)[fast_nc(NW, LT)], int (&qacc)[4], int (&qinv)[4]) {
  const uint32_t L = LT > 0 ? (uint32_t)LT : Lrt;
  constexpr int NC = fast_nc(NW, LT);
  // ---- reference window: 2-bit codes and invalid bits -------------------------------------------------
  uint32_t rf[NW], rd[NW], ve[NW];   // ref codes, read codes, valid (even bit of each position)
  bool ve_dirty;                     // some position < L is invalid (else ve is the plain length mask)
  {
    const uint32_t wi = g0 >> 4, sh = (g0 & 15u) * 2u;
    uint32_t w[NW + 1];
#pragma unroll
    for (int k = 0; k <= NW; ++k) w[k] = __ldg(ref.seq2 + wi + k);
    const uint32_t ii = g0 >> 5, s1 = g0 & 31u;
    const uint32_t i0 = __ldg(ref.inv + ii), i1 = __ldg(ref.inv + ii + 1), i2 = __ldg(ref.inv + ii + 2);
#pragma unroll
    for (int k = 0; k < NW; ++k) rf[k] = __funnelshift_r(w[k], w[k + 1], sh);
    uint32_t iv[2] = {__funnelshift_r(i0, i1, s1), __funnelshift_r(i1, i2, s1)};
    const uint32_t mask0 = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
    const uint32_t mask1 = L > 32 ? (L >= 64 ? 0xFFFFFFFFu : ((1u << (L - 32)) - 1u)) : 0u;
    const uint32_t any = (iv[0] & mask0) | (iv[1] & mask1);
#pragma unroll
    for (int k = 0; k < NW; ++k) ve[k] = lenmask[k] & 0x55555555u;
    ve_dirty = any != 0;
    if (any) {   // rare: N / IUPAC in the window -> clear those positions
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        uint32_t h = (iv[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;   // 16 invalid bits -> even bits of 32
        h = (h | (h << 8)) & 0x00FF00FFu;
        h = (h | (h << 4)) & 0x0F0F0F0Fu;
        h = (h | (h << 2)) & 0x33333333u;
        h = (h | (h << 1)) & 0x55555555u;
        ve[k] &= ~h;
      }
    }
  }
  // ---- N / IUPAC calls of this read: the warp has OR-ed them into this read's row of the invalid map --------
  if (has_n) {
    ve_dirty = true;
#pragma unroll
    for (int k = 0; k < NW; ++k) { ve[k] &= ~inv_row[k]; inv_row[k] = 0; }
  }
  // ---- read codes (unaligned row in shared memory) ----------------------------------------------------
  {
    uint32_t w[NW + 1];
#pragma unroll
    for (int k = 0; k <= NW; ++k) w[k] = brow_w[k];
#pragma unroll
    for (int k = 0; k < NW; ++k) rd[k] = __funnelshift_r(w[k], w[k + 1], bshift) & lenmask[k];
  }
  // ---- minus strand: reverse-complement both arrays (qualities stay forward, Q10) ------------------------
  if (rev) {
    const uint32_t s = 2u * (16u * NW - L);   // < 32
    uint32_t a[NW], b[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) { a[k] = __brev(rf[NW - 1 - k]); b[k] = __brev(rd[NW - 1 - k]); }
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const uint32_t an = k + 1 < NW ? a[k + 1] : 0u, bn = k + 1 < NW ? b[k + 1] : 0u;
      uint32_t x = __funnelshift_r(a[k], an, s), y = __funnelshift_r(b[k], bn, s);
      // brev swapped the two bits of every code: swap back, then complement (A<->T, C<->G is bitwise NOT)
      x = ~(((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u));
      y = ~(((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u));
      rf[k] = x & lenmask[k];
      rd[k] = y & lenmask[k];
    }
    if (ve_dirty) {   // the plain length mask is its own mirror image; only a punctured one needs reversing
      uint32_t v[NW];
#pragma unroll
      for (int k = 0; k < NW; ++k) v[k] = __brev(ve[NW - 1 - k]);
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        const uint32_t vn = k + 1 < NW ? v[k + 1] : 0u;
        ve[k] = (__funnelshift_r(v[k], vn, s) >> 1) & 0x55555555u;   // brev moved the even (valid) bit to the odd one
      }
    }
  }
  // ---- match / mismatch masks and one-hot match words ---------------------------------------------------
  uint32_t mm[NW];   // positions to visit one by one: mismatches and invalid positions (even bits)
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    const uint32_t x = rf[k] ^ rd[k];
    const uint32_t ne = (x | (x >> 1)) & 0x55555555u;
    const uint32_t m = ~ne & ve[k];
    // counted lanes per position (c1 c0 = read code of a matching base):
    //   U: even bit E0 = match,        odd bit E2 = match & c1      (G or T)
    //   V: even bit E1 = match & c0,   odd bit E3 = match & c1 & c0 (T)      (C or T)
    // A = E0-E1-E2+E3, C = E1-E3, G = E2-E3, T = E3 are formed once, at the block flush.
    const uint32_t ms = m << 1, rds = rd[k] << 1;
    const uint32_t ac = m | (ms & rd[k]);
    const uint32_t gt = rd[k] & (m | (ms & rds));
    if (k == NW - 1 && NC == 2 * NW - 1) cw[2 * k] = ac | (gt << 16);
    else { cw[2 * k] = ac; cw[2 * k + 1 < NC ? 2 * k + 1 : 0] = gt; }
    mm[k] = (lenmask[k] & 0x55555555u) & ~m;
  }
  // ---- quality sums by read base over ALL positions < L (corrected for mismatches / invalid at the end) ---
  {
    constexpr int NQ = 4 * NW;           // quality words (4 positions each)
    const int nq = (int)((L + 3) >> 2);
#pragma unroll
    for (int h = 0; h < 2 * NW; ++h) {   // 8 positions per selector word
      if (8 * h >= (int)L) break;        // L is uniform over the launch: no divergence
      uint32_t s = (rd[h >> 1] >> (16 * (h & 1))) & 0xFFFFu;
      s = (s | (s << 8)) & 0x00FF00FFu;
      s = (s | (s << 4)) & 0x0F0F0F0Fu;
      s = (s | (s << 2)) & 0x33333333u;
      const int first = 8 * h;
      if ((int)L < first + 8) s |= 0x44444444u << (4 * (L - first));   // positions >= L select a zero byte
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = 2 * h + j;
        if (k < NQ && 4 * k < (int)L) {
          uint32_t q = qrow_w[k];
          if (LT == 0 || (LT & 3)) {     // rows start at any byte unless L is a multiple of 4
            const uint32_t q1 = (k + 1 <= nq) ? qrow_w[k + 1] : 0u;
            q = __funnelshift_r(q, q1, qshift);
          }
          const uint32_t sel = j ? (s >> 16) : s;
          // qacc[0] collects the sum over all bases (A = total - C - G - T at the flush)
          qacc[0] = dp4a_ss(q, (4 * k + 4 <= (int)L) ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (8 * (4 * k + 4 - (int)L))), qacc[0]);
          qacc[1] = dp4a_ss(q, __byte_perm(tbl[1], 0u, sel), qacc[1]);
          qacc[2] = dp4a_ss(q, __byte_perm(tbl[2], 0u, sel), qacc[2]);
          qacc[3] = dp4a_ss(q, __byte_perm(tbl[3], 0u, sel), qacc[3]);
        }
      }
    }
  }
  // ---- mismatching / invalid positions, one at a time.  Words 0,1 share the even/odd bits of w0, words 2,3 of w1;
  //      one loop drains both so a warp pays max-over-lanes iterations once ------------------------------------
  {
    uint32_t w0 = mm[0] | (NW > 1 ? (mm[NW > 1 ? 1 : 0] << 1) : 0u);
    uint32_t w1 = NW > 2 ? (mm[NW > 2 ? 2 : 0] | (NW > 3 ? (mm[NW > 3 ? 3 : 0] << 1) : 0u)) : 0u;
    while (w0 | w1) {
      const bool first = w0 != 0;
      const uint32_t word = first ? w0 : w1;
      const int b = __ffs((int)word) - 1;
      if (first) w0 &= w0 - 1; else w1 &= w1 - 1;
      const int k = (first ? 0 : 2) + (b & 1);
      const int sh = b & ~1;
      uint32_t rfw = rf[0], rdw = rd[0], vew = ve[0];
#pragma unroll
      for (int j = 1; j < NW; ++j)
        if (k == j) { rfw = rf[j]; rdw = rd[j]; vew = ve[j]; }
      const uint32_t i = 16u * k + (sh >> 1);
      const uint32_t a = (rfw >> sh) & 3u, bb = (rdw >> sh) & 3u;
      const int q = (int)(signed char)qrow_b[i];
      if ((vew >> sh) & 1u) {
        atomicAdd(&F.s_mm_cnt[i * 16 + a * 4 + bb], 1u);
        atomicAdd(&F.s_mm_q[i * 16 + a * 4 + bb], (uint32_t)q);
      } else {   // invalid position: its quality went into the all-position sums, take it out again
        qinv[0] += bb == 0u ? q : 0; qinv[1] += bb == 1u ? q : 0; qinv[2] += bb == 2u ? q : 0; qinv[3] += bb == 3u ? q : 0;
      }
    }
  }
}

template <int NW, int NPL, int LT>
__global__ void __launch_bounds__(FAST_THREADS, 5) profile_fast_kernel(const ProfileParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NC = fast_nc(NW, LT);
  const uint32_t max_len = P.lay.max_len;
  const uint32_t L = LT > 0 ? (uint32_t)LT : P.b.uniform_len;
  const uint32_t bpr = (L + 3) >> 2;
  const FastStage lay = fast_stage_layout(L);
  const FastLayout fl = fast_layout(max_len, L, NW);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  FastSmem F;
  F.s_misc = reinterpret_cast<unsigned long long*>(smem_raw + fl.misc);
  F.s_mm_cnt = reinterpret_cast<uint32_t*>(smem_raw + fl.mm_cnt);
  F.s_mm_q = reinterpret_cast<uint32_t*>(smem_raw + fl.mm_q);
  F.s_fast = reinterpret_cast<uint32_t*>(smem_raw + fl.fast);
  unsigned char* wbase = smem_raw + fl.warp0 + (size_t)warp * fl.warp_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + fl.bars);
  uint32_t* s_inv = reinterpret_cast<uint32_t*>(wbase + fl.inv);        // [WT_READS][NW], zero between uses
  unsigned char* stage0 = wbase + fl.stage0;

  for (uint32_t k = threadIdx.x; k < fl.warp0 / 4; k += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[k] = 0;
  for (uint32_t k = lane; k < WT_READS * NW; k += 32) s_inv[k] = 0;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < FAST_STAGES; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint64_t n_reads = P.b.n_reads;
  const uint32_t n_wt = P.n_tiles;                       // warp-tiles in the batch (the last may be partial)
  const uint32_t n_full = (uint32_t)(n_reads / WT_READS);
  const uint32_t gw = blockIdx.x * FAST_WARPS + warp;    // global warp id
  const uint32_t GW = gridDim.x * FAST_WARPS;
  auto issue = [&](uint32_t wt, uint32_t slot) {         // lane 0 only; partial tiles are staged by the warp itself
    if (wt >= n_full) return;
    unsigned char* dst = stage0 + (size_t)slot * lay.total;
    const uint64_t r0 = (uint64_t)wt * WT_READS;
    const uint32_t bb = WT_READS * bpr, qb = WT_READS * L;
    mbar_expect_tx(&bars[slot], 3 * WT_READS * 4 + bb + qb);
    bulk_g2s(dst + lay.meta, P.b.meta + r0, WT_READS * 4, &bars[slot]);
    bulk_g2s(dst + lay.start, P.b.ref_start + r0, WT_READS * 4, &bars[slot]);
    bulk_g2s(dst + lay.cigar, P.b.cigar + r0, WT_READS * 4, &bars[slot]);
    bulk_g2s(dst + lay.bases, P.b.bases2 + r0 * bpr, bb, &bars[slot]);
    bulk_g2s(dst + lay.qual, P.b.qual + r0 * L, qb, &bars[slot]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < FAST_STAGES; ++s) {
      const uint32_t wt = gw + (uint32_t)s * GW;
      if (wt < n_wt) issue(wt, s);
    }
  }

  uint32_t lenmask[NW];
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    const int rem = (int)L - 16 * k;
    lenmask[k] = rem >= 16 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << (2 * rem)) - 1u));
  }
  // PRMT source tables (byte b = 0xFF): kept opaque so they live in registers instead of being re-materialised
  uint32_t tbl[4];
  asm volatile("mov.u32 %0, 0x000000FF;" : "=r"(tbl[0]));
  asm volatile("mov.u32 %0, 0x0000FF00;" : "=r"(tbl[1]));
  asm volatile("mov.u32 %0, 0x00FF0000;" : "=r"(tbl[2]));
  asm volatile("mov.u32 %0, 0xFF000000;" : "=r"(tbl[3]));

  Planes<NPL> pl[NC];
  uint32_t tot[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    tot[c] = 0;
#pragma unroll
    for (int p = 0; p < NPL; ++p) pl[c].p[p] = 0;
  }
  int qacc[4] = {0, 0, 0, 0}, qinv[4] = {0, 0, 0, 0};
  uint32_t n_fast = 0, since_flush = 0;
  constexpr uint32_t kFlushEvery = (1u << NPL) - 2u;   // reads a thread may add before a counter could overflow
  uint64_t c_lo = 1, c_hi = 0;                         // contig bounds cached from the previous warp-tile

  auto flush_vc = [&]() {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      tot[c] += vc_warp_total<NPL>(pl[c]);
#pragma unroll
      for (int p = 0; p < NPL; ++p) pl[c].p[p] = 0;
    }
    since_flush = 0;
    qacc[0] -= qacc[1] + qacc[2] + qacc[3];   // qacc[0] held the all-base total
#pragma unroll
    for (int b = 0; b < 4; ++b) {        // dp4a accumulated -q; keep the int32 far from overflow
      const int t = __reduce_add_sync(0xFFFFFFFFu, qacc[b]);
      const int u = __reduce_add_sync(0xFFFFFFFFu, qinv[b]);
      if (lane == 0 && t) atomicAdd(&F.s_misc[b], (unsigned long long)(long long)(-t));
      if (lane == 0 && u) atomicAdd(&F.s_misc[4 + b], (unsigned long long)(long long)u);
      qacc[b] = 0;
      qinv[b] = 0;
    }
  };

  // exception (N call) list of the first warp-tile's tile
  uint32_t e_lo = 0, e_hi = 0;
  if (gw < n_wt) { e_lo = __ldg(P.b.tile_exc_off + (gw >> 2)); e_hi = __ldg(P.b.tile_exc_off + (gw >> 2) + 1); }

  uint32_t it = 0;
  for (uint32_t wt = gw; wt < n_wt; wt += GW, ++it) {
    const uint32_t slot = it % FAST_STAGES;
    const uint32_t parity = (it / FAST_STAGES) & 1u;
    // this tile's N calls (issued before the wait so the latency overlaps it) and the next tile's list bounds
    const uint32_t ex = (e_lo + lane < e_hi) ? __ldg(P.b.exc + e_lo + lane) : 0xFFFFFFFFu;
    const uint32_t cur_lo = e_lo, cur_hi = e_hi;
    if (wt + GW < n_wt) {
      e_lo = __ldg(P.b.tile_exc_off + ((wt + GW) >> 2));
      e_hi = __ldg(P.b.tile_exc_off + ((wt + GW) >> 2) + 1);
    }
    unsigned char* sbw = stage0 + (size_t)slot * lay.total;
    uint32_t n_here = WT_READS;
    if (wt < n_full) {
      mbar_wait(&bars[slot], parity);
    } else {   // the batch's last, partial warp-tile: bounds-checked cooperative copy instead of TMA
      const uint64_t r0 = (uint64_t)wt * WT_READS;
      n_here = (uint32_t)(n_reads - r0);
      uint32_t* d32 = reinterpret_cast<uint32_t*>(sbw);
      for (uint32_t k = lane; k < WT_READS; k += 32) {
        const bool in = k < n_here;
        d32[lay.meta / 4 + k] = in ? __ldg(P.b.meta + r0 + k) : 0u;
        d32[lay.start / 4 + k] = in ? __ldg(P.b.ref_start + r0 + k) : 0u;
        d32[lay.cigar / 4 + k] = in ? __ldg(P.b.cigar + r0 + k) : 0u;
      }
      const uint32_t* gb = reinterpret_cast<const uint32_t*>(P.b.bases2 + r0 * bpr);   // r0*bpr is a multiple of 64
      for (uint32_t k = lane; k < (n_here * bpr + 3) / 4; k += 32) d32[lay.bases / 4 + k] = __ldg(gb + k);
      const uint32_t* gq = reinterpret_cast<const uint32_t*>(P.b.qual + r0 * L);
      for (uint32_t k = lane; k < (n_here * L + 3) / 4; k += 32) d32[lay.qual / 4 + k] = __ldg(gq + k);
      __syncwarp();
    }
    const unsigned char* sb = sbw;
    const uint32_t* s_meta = reinterpret_cast<const uint32_t*>(sb + lay.meta);
    const uint32_t* s_start = reinterpret_cast<const uint32_t*>(sb + lay.start);
    const uint32_t* s_cig = reinterpret_cast<const uint32_t*>(sb + lay.cigar);
    // contig bounds: reads starting and ending inside the contig of the tile's first read pass the range test
    {
      const uint64_t g = s_start[0];
      if (!(g >= c_lo && g < c_hi)) {
        if (g < P.ref.n_bases) {
          const uint32_t c = contig_of(P.ref, g);
          c_lo = __ldg(P.ref.contig_off + c);
          c_hi = __ldg(P.ref.contig_off + c + 1);
        } else { c_lo = 1; c_hi = 0; }
      }
    }
    // scatter this warp-tile's N calls into the per-read invalid map
    {
      const uint32_t quarter = wt & 3u;
      uint32_t x = ex;
      for (uint32_t e = cur_lo; e < cur_hi; e += 32) {
        if (e != cur_lo) x = (e + lane < cur_hi) ? __ldg(P.b.exc + e + lane) : 0xFFFFFFFFu;
        if (x != 0xFFFFFFFFu && ((x >> 22) & 3u) == quarter) {
          const uint32_t rit = (x >> 16) & 63u, p = x & 0xFFFFu;
          if (p < 16u * NW) atomicOr(&s_inv[rit * NW + (p >> 4)], 1u << (2u * (p & 15u)));
        }
      }
      __syncwarp();
    }

    uint32_t xw[2][NC];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t rit = lane + h * 32;
      const uint32_t meta = s_meta[rit];
      const uint32_t g0 = s_start[rit];
      const uint32_t cg = s_cig[rit];
      const uint32_t flags = PS_META_FLAGS(meta);
      const bool has_n = (flags & PS_RF_HAS_INVALID) != 0;
      const bool fast = (flags & ~(PS_RF_REVERSE | PS_RF_HAS_INVALID)) == 0 && op_is_match(cg & 15u) &&
                        (cg >> 4) == L && L <= max_len && (uint64_t)g0 >= c_lo && (uint64_t)g0 + L <= c_hi &&
                        rit < n_here;
#pragma unroll
      for (int c = 0; c < NC; ++c) xw[h][c] = 0;
      if (fast) {
        const uint32_t boff = rit * bpr, qoff = rit * L;
        fast_read<NW, LT>(P.ref, F, L, g0, (flags & PS_RF_REVERSE) != 0,
                          reinterpret_cast<const uint32_t*>(sb + lay.bases + (boff & ~3u)), (boff & 3u) * 8u,
                          reinterpret_cast<const uint32_t*>(sb + lay.qual + (qoff & ~3u)), (qoff & 3u) * 8u,
                          sb + lay.qual + qoff, lenmask, s_inv + rit * NW, has_n, tbl, xw[h], qacc, qinv);
        ++n_fast;
      } else if (rit < n_here) {
        if (has_n) {
#pragma unroll
          for (int k = 0; k < NW; ++k) s_inv[rit * NW + k] = 0;
        }
        P.deferred[atomicAdd(P.deferred_count, 1u)] = (uint32_t)(P.first_read + (uint64_t)wt * WT_READS + rit);
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) vc_add2<NPL>(pl[c].p, xw[0][c], xw[1][c]);
    since_flush += 2;
    if (since_flush >= kFlushEvery) flush_vc();

    __syncwarp();   // every lane is done with this stage buffer
    if (lane == 0) {
      const uint32_t nxt = wt + FAST_STAGES * GW;
      if (nxt < n_wt) issue(nxt, slot);
    }
  }
  flush_vc();

  // ---- per-thread totals -> shared histograms ----------------------------------------------------------------
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (tot[c] == 0) continue;
    uint32_t i, base;
    if (c == 2 * (NW - 1) && NC == 2 * NW - 1) {      // packed last word: A|C in lanes 0..15, G|T in 16..31
      i = 16u * (NW - 1) + ((lane & 15u) >> 1);
      base = 2u * (lane & 1u) + (lane >> 4);        // U half: E0/E2, V half: E1/E3
    } else {
      i = 16u * (c >> 1) + (lane >> 1);
      base = 2u * (lane & 1u) + (c & 1u);
    }
    if (i < max_len) atomicAdd(&F.s_fast[i * 4 + base], tot[c]);
  }
  {
    const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, n_fast);
    if (lane == 0 && t) atomicAdd(&F.s_misc[8], (unsigned long long)t);
  }
  __syncthreads();
  // ---- block flush ---------------------------------------------------------------------------------------------
  for (uint32_t k = threadIdx.x; k < max_len * 16; k += blockDim.x) {
    const uint32_t a = (k >> 2) & 3u, b = k & 3u, i = k >> 4;
    const unsigned long long cnt = a == b ? fast_base_count(F.s_fast + i * 4, a) : F.s_mm_cnt[k];
    if (cnt) atomicAdd(P.acc + P.lay.conv + k, cnt);
  }
  if (threadIdx.x < 16) {   // per (ref, read) pair: counts and mismatch quality over all positions
    const uint32_t a = threadIdx.x >> 2, b = threadIdx.x & 3u;
    unsigned long long cnt = 0;
    long long qs = 0;
    for (uint32_t i = 0; i < max_len; ++i) {
      cnt += a == b ? fast_base_count(F.s_fast + i * 4, a) : F.s_mm_cnt[i * 16 + threadIdx.x];
      if (a != b) qs += (long long)(int)F.s_mm_q[i * 16 + threadIdx.x];
    }
    if (cnt) {
      atomicAdd(P.acc + P.lay.qcnt + threadIdx.x, cnt);   // fast reads never hold I/D: every counted base has a quality
      atomicAdd(P.acc + P.lay.ctr + PS_PC_TOTAL_BASES_CHECKED, cnt);
    }
    if (a != b && qs) {
      atomicAdd(P.acc + P.lay.qsum + threadIdx.x, (unsigned long long)qs);
      atomicAdd(&F.s_misc[12 + b], (unsigned long long)qs);             // mismatch quality by read base
    }
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const uint32_t b = threadIdx.x;
    const long long v = (long long)F.s_misc[b] - (long long)F.s_misc[4 + b] - (long long)F.s_misc[12 + b];
    if (v) atomicAdd(P.acc + P.lay.qsum + b * 5, (unsigned long long)v);
  }
  if (threadIdx.x == 8 && F.s_misc[8]) {
    atomicAdd(P.acc + P.lay.ctr + PS_PC_NUM_READS_PROCESSED, F.s_misc[8]);
    atomicAdd(P.fault + 1, F.s_misc[8]);   // debug word: reads that took the fast path
  }
}

// a block may add this many reads before the 32-bit shared mismatch-quality cells could overflow
#define FAST_MAX_READS_PER_BLOCK (1u << 18)

template <int NW, int NPL, int LT>
cudaError_t launch_fast(ps_ctx* ctx, const ProfileParams& P, uint32_t n_wt, cudaStream_t stream) {
  const size_t smem = fast_layout(P.lay.max_len, P.b.uniform_len, NW).total + 128;
  auto kern = profile_fast_kernel<NW, NPL, LT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FAST_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  uint32_t grid = (uint32_t)ctx->sm_count * (uint32_t)per_sm;
  const uint32_t need = (n_wt + FAST_WARPS - 1) / FAST_WARPS;
  if (grid > need) grid = need;
  // keep every block under FAST_MAX_READS_PER_BLOCK by splitting very large batches into several launches
  const uint64_t per_launch = (uint64_t)grid * FAST_MAX_READS_PER_BLOCK / WT_READS;
  for (uint64_t first = 0; first < n_wt; first += per_launch) {
    ProfileParams Q = P;
    const uint64_t cnt = std::min<uint64_t>(per_launch, n_wt - first);
    Q.n_tiles = (uint32_t)cnt;
    const uint64_t r0 = first * WT_READS;
    Q.b.meta += r0; Q.b.ref_start += r0; Q.b.cigar += r0;
    Q.b.bases2 += r0 * ((P.b.uniform_len + 3) / 4);
    Q.b.qual += r0 * P.b.uniform_len;
    Q.b.tile_exc_off += r0 / PS_TILE_READS;
    Q.b.n_reads = std::min<uint64_t>(P.b.n_reads - r0, cnt * WT_READS);
    Q.first_read = r0;    // offset added to deferred read indices
    kern<<<grid, FAST_THREADS, smem, stream>>>(Q);
    ctx->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
