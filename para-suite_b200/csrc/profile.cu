// K1: error-profile count kernels.  Replace the loop ErrorProfiling.java:146-409 (reference:
// /root/reference/src/src/utils/errorprofile/ErrorProfiling.java) for one SoA batch.
//
// profile_fast_kernel   the common PAR-CLIP shape: uniform read length L <= 64, one M/=/X cigar op (L == R).
//   * persistent CTAs; each iteration stages a 512-read super-tile (meta, ref_start, cigar, 2-bit bases,
//     qualities: contiguous runs in HBM) into shared memory with cp.async.bulk + mbarrier, 3 stages deep;
//   * one thread per read, all per-base work bit-parallel on 2-bit packed words:
//       match counts   -> per-thread bit-sliced ("vertical") counters, one-hot (A|C, G|T) x position lanes,
//                         merged across the warp with a carry-save butterfly every 2^P-2 reads; no atomics
//       mismatches     -> rare: one 64-bit shared atomic (count | quality sum) per mismatching base
//       quality sums   -> dp4a of the quality bytes against 0/-1 byte masks built with PRMT from the codes
//   * reads that do not fit the fast shape (flags, N calls, other cigars, contig edges) fall through to the
//     generic per-read routine inside the same kernel.
// profile_generic_kernel  every CIGAR / every flag; literal per-read walk (also the tail of a batch).
//
// Counts land in per-block shared-memory histograms and are flushed once per block with 64-bit global atomics
// into the accumulator vector (internal.h ProfileLayout) -- the unit of the multi-GPU all-reduce.
#include "device_common.cuh"

namespace {

struct ProfileParams {
  DeviceBatch b;
  DeviceRef ref;
  ProfileLayout lay;
  unsigned long long* acc;    // int64 accumulators (two's complement adds)
  unsigned long long* fault;
  uint64_t ordinal0;
  uint64_t first_read;        // generic kernel: first read it covers
  uint32_t n_tiles;           // generic kernel: tiles from first_read on;  fast kernel: number of super-tiles
};

// shared-memory histograms of the generic path
struct GenericSmem {
  unsigned long long* s_q;    // [32] qualityPerMismatch sums | counts
  unsigned long long* s_ctr;  // [8]
  uint32_t* s_conv;           // [max_len*16]
};

// ---------------------------------------------------------------------------------------------------------
// Generic per-read routine: literal restatement of ErrorProfiling.java:155-408 on packed data.
// ---------------------------------------------------------------------------------------------------------
__device__ __noinline__ void profile_read_generic(const ProfileParams& P, const GenericSmem& S, uint64_t tile,
                                                  uint32_t rit, uint64_t r, uint32_t meta, ReadOffsets off) {
  const uint32_t max_len = P.lay.max_len;
  const uint32_t flags = PS_META_FLAGS(meta), L = PS_META_LEN(meta), ncig = PS_META_NCIGAR(meta);
  const uint64_t ordinal = P.ordinal0 + r;
  // filters :155-166 (first match wins)
  if (flags & PS_RF_UNMAPPED) { atomicAdd(&S.s_ctr[PS_PC_UNMAPPED], 1ull); return; }
  if (flags & PS_RF_DUPLICATE) { atomicAdd(&S.s_ctr[PS_PC_DUPLICATES], 1ull); return; }
  if (flags & PS_RF_POS_ZERO) { atomicAdd(&S.s_ctr[PS_PC_START_ZERO], 1ull); return; }

  const uint32_t* cig = P.b.cigar + off.cigar;
  uint32_t R = 0;
  bool has_indel = false;
  for (uint32_t e = 0; e < ncig; ++e) {
    uint32_t c = __ldg(cig + e), op = c & 15u;
    if (op_consumes_ref(op)) R += c >> 4;
    has_indel |= (op == 1u) | (op == 2u);
  }
  const uint64_t g0 = __ldg(P.b.ref_start + r);
  {  // :169-172 FASTA fetch range
    bool bad = (flags & PS_RF_REF_RANGE) || g0 >= P.ref.n_bases;
    if (!bad) {
      uint32_t c = contig_of(P.ref, g0);
      bad = g0 + R > __ldg(P.ref.contig_off + c + 1);
    }
    if (bad) { raise_fault(P.fault, ordinal, PS_THROW_REF_RANGE); return; }
  }
  atomicAdd(&S.s_ctr[PS_PC_NUM_READS_PROCESSED], 1ull);                        // :174
  if (R == 0) { raise_fault(P.fault, ordinal, PS_THROW_EMPTY_REF); return; }

  const uint32_t ml = L > R ? L : R;
  const bool walked = L != R;
  bool skip = false;
  if (walked) {   // :194-299, pass 1: bounds (skip), indel side effects, uncaught exceptions
    int64_t pr = 0, pq = 0, pm = 0;
    for (uint32_t e = 0; e < ncig; ++e) {
      const uint32_t c = __ldg(cig + e), op = c & 15u;
      const int64_t n = c >> 4;
      if (op_is_match(op)) {
        // any z with z+pm >= ml, z+pr >= R or z+pq >= L throws inside the try -> skip
        if (n > 0 && (pm + n > (int64_t)ml || pr + n > (int64_t)R || pq + n > (int64_t)L)) skip = true;
        pm += n; pr += n; pq += n;
      } else if (op == 3u) {
        pr += n; pq += n;
      } else if (op == 1u || op == 2u) {
        if (n > 0 && pm + n > (int64_t)ml) { raise_fault(P.fault, ordinal, PS_THROW_INDEL_FILL); return; }
        pm += n;
        if (op == 1u) pq += n; else pr += n;
        unsigned long long* arr = P.acc + (op == 1u ? P.lay.ins : P.lay.del);
        for (int64_t q = 1; q <= n; ++q) {
          if (pm + q >= (int64_t)max_len) { raise_fault(P.fault, ordinal, PS_THROW_INDEL_POS); return; }
          atomicAdd(arr + (pm + q), 1ull);
        }
        if (n > 1) atomicAdd(&S.s_ctr[PS_PC_LONGER_INDELS], 1ull);
      }
    }
    atomicAdd(&S.s_ctr[PS_PC_INDEL_READ], 1ull);                               // :296
  }
  if (skip) { atomicAdd(&S.s_ctr[PS_PC_SKIPPED_READS], 1ull); return; }        // :303-306

  // count loop :349-408 over the (virtual) temp arrays
  const bool rev = flags & PS_RF_REVERSE;
  const bool has_inv = flags & PS_RF_HAS_INVALID;
  const uint32_t qual_len = (flags & PS_RF_QUAL_MISSING) ? 0u : L;
  const uint8_t* rb = P.b.bases2 + off.base;
  const uint8_t* rq = P.b.qual + off.qual;
  long long q_acc0 = 0, q_acc1 = 0, q_acc2 = 0, q_acc3 = 0;
  uint32_t q_cnt0 = 0, q_cnt1 = 0, q_cnt2 = 0, q_cnt3 = 0;
  uint32_t checked = 0;
  uint32_t f_i = 0xFFFFFFFFu, f_code = 0;   // first uncaught exception of the count loop (by position i)
  bool live = true;
  // forward strand: columns ascend with the cigar; reverse strand: i = ml-1-col, so iterate ops backwards.
  const uint32_t n_ops = walked ? ncig : 1u;
  for (uint32_t step = 0; step < n_ops && live; ++step) {
    int64_t seg_col, seg_ref, seg_read, seg_len;
    if (!walked) {
      seg_col = 0; seg_ref = 0; seg_read = 0; seg_len = L;   // L == R: ungapped compare (Q1)
    } else {
      const uint32_t e_target = rev ? ncig - 1 - step : step;
      int64_t pr = 0, pq = 0, pm = 0;
      uint32_t c = 0;
      for (uint32_t e = 0; e <= e_target; ++e) {
        c = __ldg(cig + e);
        if (e == e_target) break;
        const uint32_t op = c & 15u;
        const int64_t n = c >> 4;
        if (op_is_match(op)) { pm += n; pr += n; pq += n; }
        else if (op == 3u) { pr += n; pq += n; }
        else if (op == 1u) { pm += n; pq += n; }
        else if (op == 2u) { pm += n; pr += n; }
      }
      if (!op_is_match(c & 15u)) continue;
      seg_col = pm; seg_ref = pr; seg_read = pq; seg_len = c >> 4;
    }
    for (int64_t zz = 0; zz < seg_len; ++zz) {
      const int64_t z = rev ? seg_len - 1 - zz : zz;
      const int64_t col = seg_col + z;
      const uint32_t i = (uint32_t)(rev ? (int64_t)ml - 1 - col : col);
      const uint64_t g = g0 + (uint64_t)(seg_ref + z);
      const uint32_t p = (uint32_t)(seg_read + z);
      bool ok = !ref_invalid_at(P.ref, g);
      if (ok && has_inv) ok = !read_pos_invalid(P.b, tile, rit, p);
      if (!ok) continue;
      uint32_t a = ref_code_at(P.ref, g), b = read_code_at(rb, p);
      if (rev) { a = 3u - a; b = 3u - b; }
      if (i >= max_len) { f_i = i; f_code = PS_THROW_POS_MAXLEN; live = false; break; }
      atomicAdd(&S.s_conv[i * 16 + a * 4 + b], 1u);
      ++checked;
      if (!has_indel) {
        if (i >= qual_len) { f_i = i; f_code = PS_THROW_QUAL_RANGE; live = false; break; }
        const long long qv = (long long)(signed char)__ldg(rq + i);   // qualities are NOT reversed (Q10)
        if (a == b) {
          if (a == 0) { q_acc0 += qv; q_cnt0++; } else if (a == 1) { q_acc1 += qv; q_cnt1++; }
          else if (a == 2) { q_acc2 += qv; q_cnt2++; } else { q_acc3 += qv; q_cnt3++; }
        } else {
          atomicAdd(&S.s_q[a * 4 + b], (unsigned long long)qv);
          atomicAdd(&S.s_q[16 + a * 4 + b], 1ull);
        }
      }
    }
  }
  if (P.lay.infer_q) {   // :402-407 touches baseQualitiesPerPos[i] / readQualities[i] for EVERY i < ml
    const uint32_t iq = max_len < qual_len ? max_len : qual_len;
    if (ml > iq && iq < f_i) {
      f_i = iq;
      f_code = max_len <= qual_len ? PS_THROW_POS_MAXLEN : PS_THROW_QUAL_RANGE;
      live = false;
    }
  }
  if (f_i != 0xFFFFFFFFu) raise_fault(P.fault, ordinal, f_code);
  if (!live) return;
  if (q_cnt0) { atomicAdd(&S.s_q[0], (unsigned long long)q_acc0); atomicAdd(&S.s_q[16], (unsigned long long)q_cnt0); }
  if (q_cnt1) { atomicAdd(&S.s_q[5], (unsigned long long)q_acc1); atomicAdd(&S.s_q[21], (unsigned long long)q_cnt1); }
  if (q_cnt2) { atomicAdd(&S.s_q[10], (unsigned long long)q_acc2); atomicAdd(&S.s_q[26], (unsigned long long)q_cnt2); }
  if (q_cnt3) { atomicAdd(&S.s_q[15], (unsigned long long)q_acc3); atomicAdd(&S.s_q[31], (unsigned long long)q_cnt3); }
  atomicAdd(&S.s_ctr[PS_PC_TOTAL_BASES_CHECKED], (unsigned long long)checked);
  if (P.lay.infer_q)
    for (uint32_t i = 0; i < ml; ++i) atomicAdd(P.acc + P.lay.qhist + (size_t)i * 256 + __ldg(rq + i), 1ull);
}

__device__ __forceinline__ void flush_generic(const ProfileParams& P, const GenericSmem& S) {
  for (uint32_t k = threadIdx.x; k < P.lay.max_len * 16; k += blockDim.x)
    if (S.s_conv[k]) atomicAdd(P.acc + P.lay.conv + k, (unsigned long long)S.s_conv[k]);
  if (threadIdx.x < 32 && S.s_q[threadIdx.x]) atomicAdd(P.acc + P.lay.qsum + threadIdx.x, S.s_q[threadIdx.x]);
  if (threadIdx.x < PS_PC_COUNT && S.s_ctr[threadIdx.x])
    atomicAdd(P.acc + P.lay.ctr + threadIdx.x, S.s_ctr[threadIdx.x]);
}

// Shared memory: s_q[32] u64 | s_ctr[8] u64 | scan scratch[8] u64 | s_conv[max_len*16] u32
__global__ void __launch_bounds__(PS_BLOCK_THREADS) profile_generic_kernel(const ProfileParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenericSmem S;
  S.s_q = reinterpret_cast<unsigned long long*>(smem_raw);
  S.s_ctr = S.s_q + 32;
  uint64_t* s_scan = reinterpret_cast<uint64_t*>(S.s_ctr + 8);
  S.s_conv = reinterpret_cast<uint32_t*>(s_scan + 8);
  for (uint32_t k = threadIdx.x; k < P.lay.max_len * 16; k += blockDim.x) S.s_conv[k] = 0;
  if (threadIdx.x < 48) S.s_q[threadIdx.x] = 0;   // s_q, s_ctr, s_scan are contiguous
  __syncthreads();
  const uint64_t tile0 = P.first_read / PS_TILE_READS;   // first_read is tile aligned
  for (uint32_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
    const uint64_t tile = tile0 + t;
    const uint64_t r = tile * PS_TILE_READS + threadIdx.x;
    const bool in_range = r < P.b.n_reads;
    const uint32_t meta = in_range ? __ldg(P.b.meta + r) : 0;
    const ReadOffsets off = read_offsets(P.b, tile, r, meta, in_range, s_scan);
    if (in_range) profile_read_generic(P, S, tile, threadIdx.x, r, meta, off);
  }
  __syncthreads();
  flush_generic(P, S);
}

// ---------------------------------------------------------------------------------------------------------
// Fast path
// ---------------------------------------------------------------------------------------------------------
#define FAST_STAGES 3
#define FAST_READS 512          // reads per super-tile (2 per thread)

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

struct FastStage {      // byte offsets inside one stage buffer
  uint32_t meta, start, cigar, bases, qual, total;
};
__host__ __device__ inline FastStage fast_stage_layout(uint32_t L) {
  FastStage s;
  const uint32_t bpr = (L + 3) / 4;
  s.meta = 0;
  s.start = FAST_READS * 4;
  s.cigar = 2 * FAST_READS * 4;
  s.bases = 3 * FAST_READS * 4;
  s.qual = s.bases + ((FAST_READS * bpr + 15) & ~15u) + 16;   // +16: word reads may run past a row end
  s.total = (s.qual + FAST_READS * L + 16 + 127) & ~127u;
  return s;
}

// bit-sliced counter: planes[p] holds bit p of 32 independent counters
template <int NPL>
__device__ __forceinline__ void vc_add2(uint32_t (&pl)[NPL], uint32_t x, uint32_t y) {
  // full adder into plane 0, ripple the carry up
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

// sum the bit-sliced counters of the 32 lanes; lane l ends up with the integer total of bit-lane l
template <int NPL>
__device__ __forceinline__ uint32_t vc_warp_total(uint32_t (&pl)[NPL]) {
  uint32_t a[NPL + 5];
#pragma unroll
  for (int p = 0; p < NPL; ++p) a[p] = pl[p];
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
  unsigned long long* s_mm;     // [max_len*16] mismatches: quality sum << 32 | count
  unsigned long long* s_misc;   // [0..3] S_all by read base, [4..7] quality at invalid positions by read base, [8] fast reads
  uint32_t* s_fast;             // [max_len*4] match counts by (position, base)
};

// One read of the fast shape.  Returns the one-hot match words for the caller's bit-sliced counters.
template <int NW>
__device__ __forceinline__ void fast_read(const ProfileParams& P, const FastSmem& F, uint32_t L, uint32_t g0, bool rev,
                                          const uint32_t* __restrict__ brow_w, uint32_t bshift,
                                          const uint32_t* __restrict__ qrow_w, uint32_t qshift,
                                          const unsigned char* __restrict__ qrow_b, const uint32_t (&lenmask)[NW],
                                          uint32_t (&ac)[NW], uint32_t (&gt)[NW], int (&qacc)[4]) {
  // ---- reference window: 2-bit codes and invalid bits -------------------------------------------------
  uint32_t rf[NW], rd[NW], ve[NW];   // ref codes, read codes, valid (even bit of each position)
  {
    const uint32_t wi = g0 >> 4, sh = (g0 & 15u) * 2u;
    uint32_t w[NW + 1];
#pragma unroll
    for (int k = 0; k <= NW; ++k) w[k] = __ldg(P.ref.seq2 + wi + k);
#pragma unroll
    for (int k = 0; k < NW; ++k) rf[k] = __funnelshift_r(w[k], w[k + 1], sh);
    const uint32_t ii = g0 >> 5, s1 = g0 & 31u;
    const uint32_t i0 = __ldg(P.ref.inv + ii), i1 = __ldg(P.ref.inv + ii + 1), i2 = __ldg(P.ref.inv + ii + 2);
    uint32_t iv[2] = {__funnelshift_r(i0, i1, s1), __funnelshift_r(i1, i2, s1)};
    const uint32_t mask0 = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
    const uint32_t mask1 = L > 32 ? (L >= 64 ? 0xFFFFFFFFu : ((1u << (L - 32)) - 1u)) : 0u;
    const uint32_t any = (iv[0] & mask0) | (iv[1] & mask1);
#pragma unroll
    for (int k = 0; k < NW; ++k) ve[k] = lenmask[k] & 0x55555555u;
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
    uint32_t a[NW], b[NW], v[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) { a[k] = __brev(rf[NW - 1 - k]); b[k] = __brev(rd[NW - 1 - k]); v[k] = __brev(ve[NW - 1 - k]); }
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const uint32_t an = k + 1 < NW ? a[k + 1] : 0u, bn = k + 1 < NW ? b[k + 1] : 0u, vn = k + 1 < NW ? v[k + 1] : 0u;
      uint32_t x = __funnelshift_r(a[k], an, s), y = __funnelshift_r(b[k], bn, s), z = __funnelshift_r(v[k], vn, s);
      // brev swapped the two bits of every code: swap back, then complement (A<->T, C<->G is bitwise NOT)
      x = ~(((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u));
      y = ~(((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u));
      rf[k] = x & lenmask[k];
      rd[k] = y & lenmask[k];
      ve[k] = (z >> 1) & 0x55555555u;   // the valid bit sat on the even bit: brev moved it to the odd one
    }
  }
  // ---- match / mismatch masks and one-hot match words ---------------------------------------------------
  uint32_t mm[NW];   // positions to visit one by one: mismatches and invalid positions (even bits)
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    const uint32_t x = rf[k] ^ rd[k];
    const uint32_t ne = (x | (x >> 1)) & 0x55555555u;
    const uint32_t m = ~ne & ve[k];
    const uint32_t lo = rd[k] & 0x55555555u, hi = (rd[k] >> 1) & 0x55555555u;
    ac[k] = (m & ~hi & ~lo) | ((m & ~hi & lo) << 1);
    gt[k] = (m & hi & ~lo) | ((m & hi & lo) << 1);
    mm[k] = (lenmask[k] & 0x55555555u) & ~m;
  }
  // ---- quality sums by read base over ALL positions < L (corrected for mismatches / invalid below) ------
  {
    constexpr int NQ = 4 * NW;           // quality words (4 positions each), those past L masked by selector
    uint32_t qw[NQ + 1];
    const int nq = (int)((L + 3) >> 2);
#pragma unroll
    for (int k = 0; k <= NQ; ++k) qw[k] = (k <= nq) ? qrow_w[k] : 0u;
#pragma unroll
    for (int h = 0; h < 2 * NW; ++h) {   // 8 positions per selector word
      if (8 * h >= (int)L) break;        // L is uniform over the launch: no divergence
      uint32_t s = (rd[h >> 1] >> (16 * (h & 1))) & 0xFFFFu;
      s = (s | (s << 8)) & 0x00FF00FFu;
      s = (s | (s << 4)) & 0x0F0F0F0Fu;
      s = (s | (s << 2)) & 0x33333333u;
      // positions >= L select a zero byte (nibble bit 2 set)
      const int first = 8 * h;
      uint32_t tail = 0;
      if ((int)L < first + 8) tail = (int)L <= first ? 0x44444444u : (0x44444444u << (4 * (L - first)));
      s |= tail;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = 2 * h + j;
        if (k < NQ && 4 * k < (int)L) {
          const uint32_t q = __funnelshift_r(qw[k], qw[k + 1], qshift);
          const uint32_t sel = j ? (s >> 16) : s;
          qacc[0] = dp4a_ss(q, __byte_perm(0x000000FFu, 0u, sel), qacc[0]);
          qacc[1] = dp4a_ss(q, __byte_perm(0x0000FF00u, 0u, sel), qacc[1]);
          qacc[2] = dp4a_ss(q, __byte_perm(0x00FF0000u, 0u, sel), qacc[2]);
          qacc[3] = dp4a_ss(q, __byte_perm(0xFF000000u, 0u, sel), qacc[3]);
        }
      }
    }
  }
  // ---- mismatching / invalid positions, one at a time -----------------------------------------------------
  uint32_t anymm = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) anymm |= mm[k];
  while (anymm) {
    int k = 0;
    uint32_t word = mm[0];
#pragma unroll
    for (int j = 1; j < NW; ++j)
      if (word == 0) { word = mm[j]; k = j; }
    const int b = __ffs((int)word) - 1;
    const uint32_t bit = 1u << b;
    uint32_t rfw = rf[0], rdw = rd[0], vew = ve[0];
#pragma unroll
    for (int j = 1; j < NW; ++j)
      if (k == j) { rfw = rf[j]; rdw = rd[j]; vew = ve[j]; }
#pragma unroll
    for (int j = 0; j < NW; ++j)
      if (k == j) mm[j] &= ~bit;
    const uint32_t i = 16u * k + (b >> 1);
    const uint32_t a = (rfw >> b) & 3u, bb = (rdw >> b) & 3u;
    const long long q = (long long)(signed char)qrow_b[i];
    if (vew & bit) atomicAdd(&F.s_mm[i * 16 + a * 4 + bb], (unsigned long long)((q << 32) + 1));
    else atomicAdd(&F.s_misc[4 + bb], (unsigned long long)q);
    anymm = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) anymm |= mm[j];
  }
}

template <int NW, int NPL>
__global__ void __launch_bounds__(PS_BLOCK_THREADS, 2) profile_fast_kernel(const ProfileParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t max_len = P.lay.max_len;
  const uint32_t L = P.b.uniform_len;
  const uint32_t bpr = (L + 3) >> 2;
  const FastStage lay = fast_stage_layout(L);
  // shared memory carve-up
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                       // [FAST_STAGES]
  GenericSmem S;
  S.s_q = reinterpret_cast<unsigned long long*>(smem_raw + 64);                 // [32]
  S.s_ctr = S.s_q + 32;                                                         // [8]
  FastSmem F;
  F.s_misc = S.s_ctr + 8;                                                       // [16]
  F.s_mm = F.s_misc + 16;                                                       // [max_len*16]
  S.s_conv = reinterpret_cast<uint32_t*>(F.s_mm + (size_t)max_len * 16);        // [max_len*16]
  F.s_fast = S.s_conv + (size_t)max_len * 16;                                   // [max_len*4]
  unsigned char* stage0 =
      smem_raw + ((64 + (32 + 8 + 16) * 8 + (size_t)max_len * (16 * 8 + 16 * 4 + 4 * 4) + 127) & ~(size_t)127);

  for (uint32_t k = threadIdx.x; k < (32 + 8 + 16) * 2 + max_len * 16 * 2; k += blockDim.x)
    reinterpret_cast<uint32_t*>(S.s_q)[k] = 0;   // s_q, s_ctr, s_misc, s_mm (u64 each)
  for (uint32_t k = threadIdx.x; k < max_len * 20; k += blockDim.x) S.s_conv[k] = 0;   // s_conv + s_fast
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < FAST_STAGES; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint32_t n_super = P.n_tiles;
  auto issue = [&](uint32_t st_idx, uint32_t slot) {   // one thread
    unsigned char* dst = stage0 + (size_t)slot * lay.total;
    const uint64_t r0 = (uint64_t)st_idx * FAST_READS;
    const uint32_t bb = FAST_READS * bpr, qb = FAST_READS * L;
    mbar_expect_tx(&bars[slot], 3 * FAST_READS * 4 + bb + qb);
    bulk_g2s(dst + lay.meta, P.b.meta + r0, FAST_READS * 4, &bars[slot]);
    bulk_g2s(dst + lay.start, P.b.ref_start + r0, FAST_READS * 4, &bars[slot]);
    bulk_g2s(dst + lay.cigar, P.b.cigar + r0, FAST_READS * 4, &bars[slot]);
    bulk_g2s(dst + lay.bases, P.b.bases2 + r0 * bpr, bb, &bars[slot]);
    bulk_g2s(dst + lay.qual, P.b.qual + r0 * L, qb, &bars[slot]);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < FAST_STAGES; ++s) {
      const uint32_t st = blockIdx.x + (uint32_t)s * gridDim.x;
      if (st < n_super) issue(st, s);
    }
  }

  // per-thread constants
  uint32_t lenmask[NW];
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    const int rem = (int)L - 16 * k;
    lenmask[k] = rem >= 16 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << (2 * rem)) - 1u));
  }
  // contig window of the first read of each super-tile bounds the cheap range test
  uint32_t ac_pl[NW][NPL], gt_pl[NW][NPL];
  uint32_t ac_tot[NW], gt_tot[NW];
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    ac_tot[k] = gt_tot[k] = 0;
#pragma unroll
    for (int p = 0; p < NPL; ++p) ac_pl[k][p] = gt_pl[k][p] = 0;
  }
  int qacc[4] = {0, 0, 0, 0};
  uint32_t n_fast = 0, since_flush = 0;
  constexpr uint32_t kFlushEvery = (1u << NPL) - 2u;   // reads a thread may add before a counter could overflow

  auto flush_vc = [&]() {
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      ac_tot[k] += vc_warp_total<NPL>(ac_pl[k]);
      gt_tot[k] += vc_warp_total<NPL>(gt_pl[k]);
#pragma unroll
      for (int p = 0; p < NPL; ++p) ac_pl[k][p] = gt_pl[k][p] = 0;
    }
    since_flush = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)          // dp4a accumulated -q; keep the int32 far from overflow
      if (qacc[b]) { atomicAdd(&F.s_misc[b], (unsigned long long)(long long)(-qacc[b])); qacc[b] = 0; }
  };

  uint32_t it = 0;
  for (uint32_t st = blockIdx.x; st < n_super; st += gridDim.x, ++it) {
    const uint32_t slot = it % FAST_STAGES;
    const uint32_t parity = (it / FAST_STAGES) & 1u;
    mbar_wait(&bars[slot], parity);
    const unsigned char* sb = stage0 + (size_t)slot * lay.total;
    const uint32_t* s_meta = reinterpret_cast<const uint32_t*>(sb + lay.meta);
    const uint32_t* s_start = reinterpret_cast<const uint32_t*>(sb + lay.start);
    const uint32_t* s_cig = reinterpret_cast<const uint32_t*>(sb + lay.cigar);
    // contig bounds of the tile's first read: reads inside it and ending inside it pass the range test
    __shared__ uint64_t s_cb[2];
    if (threadIdx.x == 0) {
      const uint64_t g = s_start[0];
      if (g < P.ref.n_bases) {
        const uint32_t c = contig_of(P.ref, g);
        s_cb[0] = __ldg(P.ref.contig_off + c);
        s_cb[1] = __ldg(P.ref.contig_off + c + 1);
      } else { s_cb[0] = 1; s_cb[1] = 0; }
    }
    __syncthreads();
    const uint64_t c_lo = s_cb[0], c_hi = s_cb[1];

    uint32_t xa[2][NW], xg[2][NW];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t rit = threadIdx.x + h * PS_BLOCK_THREADS;
      const uint64_t r = (uint64_t)st * FAST_READS + rit;
      const uint32_t meta = s_meta[rit];
      const uint32_t g0 = s_start[rit];
      const uint32_t cg = s_cig[rit];
      const uint32_t flags = PS_META_FLAGS(meta);
      const bool fast = (flags & ~PS_RF_REVERSE) == 0 && op_is_match(cg & 15u) && (cg >> 4) == L && L <= max_len &&
                        (uint64_t)g0 >= c_lo && (uint64_t)g0 + L <= c_hi && !P.lay.infer_q;
#pragma unroll
      for (int k = 0; k < NW; ++k) xa[h][k] = xg[h][k] = 0;
      if (fast) {
        const uint32_t boff = rit * bpr, qoff = rit * L;
        fast_read<NW>(P, F, L, g0, (flags & PS_RF_REVERSE) != 0,
                      reinterpret_cast<const uint32_t*>(sb + lay.bases + (boff & ~3u)), (boff & 3u) * 8u,
                      reinterpret_cast<const uint32_t*>(sb + lay.qual + (qoff & ~3u)), (qoff & 3u) * 8u,
                      sb + lay.qual + qoff, lenmask, xa[h], xg[h], qacc);
        ++n_fast;
      } else {
        ReadOffsets off;
        off.base = r * (uint64_t)bpr;
        off.qual = r * (uint64_t)L;
        off.cigar = r;
        profile_read_generic(P, S, r / PS_TILE_READS, (uint32_t)(r % PS_TILE_READS), r, meta, off);
      }
    }
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      vc_add2<NPL>(ac_pl[k], xa[0][k], xa[1][k]);
      vc_add2<NPL>(gt_pl[k], xg[0][k], xg[1][k]);
    }
    since_flush += 2;
    if (since_flush >= kFlushEvery) flush_vc();

    __syncthreads();   // everyone is done with this stage buffer
    if (threadIdx.x == 0) {
      const uint32_t nxt = st + FAST_STAGES * gridDim.x;
      if (nxt < n_super) issue(nxt, slot);
    }
  }
  flush_vc();

  // ---- per-thread totals -> shared histograms ----------------------------------------------------------------
  {
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const uint32_t i = 16u * k + (lane >> 1);
      if (i < max_len) {
        if (ac_tot[k]) atomicAdd(&F.s_fast[i * 4 + (lane & 1u)], ac_tot[k]);
        if (gt_tot[k]) atomicAdd(&F.s_fast[i * 4 + 2 + (lane & 1u)], gt_tot[k]);
      }
    }
    if (n_fast) atomicAdd(&F.s_misc[8], (unsigned long long)n_fast);
  }
  __syncthreads();
  // ---- block flush: fast-path histograms, then the generic ones -------------------------------------------------
  for (uint32_t k = threadIdx.x; k < max_len * 16; k += blockDim.x) {
    const uint32_t a = (k >> 2) & 3u, b = k & 3u, i = k >> 4;
    unsigned long long cnt = (uint32_t)F.s_mm[k];                       // mismatches (and nothing on the diagonal)
    if (a == b) cnt = F.s_fast[i * 4 + a];
    if (cnt) {
      atomicAdd(P.acc + P.lay.conv + k, cnt);
      atomicAdd(P.acc + P.lay.qcnt + (k & 15u), cnt);                   // fast reads never hold I/D: every count has a quality
      atomicAdd(P.acc + P.lay.ctr + PS_PC_TOTAL_BASES_CHECKED, cnt);
    }
    if (a != b) {
      const long long qs = (long long)F.s_mm[k] >> 32;                  // arithmetic shift: signed quality sum
      if (qs) {
        atomicAdd(P.acc + P.lay.qsum + (k & 15u), (unsigned long long)qs);
        atomicAdd(&F.s_misc[12 + b], (unsigned long long)qs);           // mismatch quality by read base
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const uint32_t b = threadIdx.x;
    const long long v = (long long)F.s_misc[b] - (long long)F.s_misc[4 + b] - (long long)F.s_misc[12 + b];
    if (v) atomicAdd(P.acc + P.lay.qsum + b * 5, (unsigned long long)v);
  }
  if (threadIdx.x == 8 && F.s_misc[8]) atomicAdd(P.acc + P.lay.ctr + PS_PC_NUM_READS_PROCESSED, F.s_misc[8]);
  flush_generic(P, S);
}

size_t fast_smem_bytes(uint32_t max_len, uint32_t L) {
  size_t head = (64 + (32 + 8 + 16) * 8 + (size_t)max_len * (16 * 8 + 16 * 4 + 4 * 4) + 127) & ~(size_t)127;
  return head + (size_t)FAST_STAGES * fast_stage_layout(L).total + 128;
}

template <int NW, int NPL>
cudaError_t launch_fast(ps_ctx* ctx, const ProfileParams& P, uint32_t n_super, cudaStream_t stream) {
  const size_t smem = fast_smem_bytes(P.lay.max_len, P.b.uniform_len);
  auto kern = profile_fast_kernel<NW, NPL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PS_BLOCK_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  uint32_t grid = (uint32_t)ctx->sm_count * (uint32_t)per_sm;
  if (grid > n_super) grid = n_super;
  ProfileParams Q = P;
  Q.n_tiles = n_super;
  kern<<<grid, PS_BLOCK_THREADS, smem, stream>>>(Q);
  ctx->launches++;
  return cudaGetLastError();
}

}  // namespace

static cudaError_t launch_generic(ps_ctx* ctx, ProfileParams P, uint64_t first_read, cudaStream_t stream) {
  if (first_read >= P.b.n_reads) return cudaSuccess;
  P.first_read = first_read;
  P.n_tiles = (uint32_t)((P.b.n_reads - first_read + PS_TILE_READS - 1) / PS_TILE_READS);
  size_t smem = 48 * 8 + (size_t)ctx->layout.max_len * 16 * 4;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(profile_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, profile_generic_kernel, PS_BLOCK_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  uint32_t grid = (uint32_t)ctx->sm_count * (uint32_t)per_sm;
  if (grid > P.n_tiles) grid = P.n_tiles;
  profile_generic_kernel<<<grid, PS_BLOCK_THREADS, smem, stream>>>(P);
  ctx->launches++;
  return cudaGetLastError();
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

cudaError_t launch_profile(ps_ctx* ctx, const DeviceBatch& b, uint64_t ordinal0, cudaStream_t stream) {
  if (b.n_reads == 0) return cudaSuccess;
  ProfileParams P;
  P.b = b;
  P.ref = ctx->ref;
  P.lay = ctx->layout;
  P.acc = static_cast<unsigned long long*>(ctx->acc.p);
  P.fault = static_cast<unsigned long long*>(ctx->fault.p);
  P.ordinal0 = ordinal0;
  P.first_read = 0;
  P.n_tiles = 0;
  uint64_t done = 0;
  const uint32_t L = b.uniform_len;
  const bool fast_ok = L >= 1 && L <= 64 && b.uniform_ncigar == 1 && !ctx->layout.infer_q && ctx->layout.max_len <= 256 &&
                       aligned16(b.meta) && aligned16(b.ref_start) && aligned16(b.cigar) && aligned16(b.bases2) &&
                       aligned16(b.qual) && b.n_reads >= FAST_READS;
  if (fast_ok) {
    const uint32_t n_super = (uint32_t)(b.n_reads / FAST_READS);
    const uint32_t nw = (L + 15) / 16;
    cudaError_t e;
    switch (nw) {
      case 1: e = launch_fast<1, 6>(ctx, P, n_super, stream); break;
      case 2: e = launch_fast<2, 6>(ctx, P, n_super, stream); break;
      case 3: e = launch_fast<3, 5>(ctx, P, n_super, stream); break;
      default: e = launch_fast<4, 5>(ctx, P, n_super, stream); break;
    }
    if (e != cudaSuccess) return e;
    done = (uint64_t)n_super * FAST_READS;
  }
  return launch_generic(ctx, P, done, stream);
}
