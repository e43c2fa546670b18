// K2/K3: T>C conversion pileup.  Replaces the loop PileupClusters.java:137-500 and
// calculateClusterInformation :585-673 (reference: /root/reference/src/src/utils/pileupclusters/).
//
// Two kernels over one coordinate-sorted SoA batch, every record assembled on the device:
//
//   pl_scan_kernel   ONE pass over the reads (tiles of 1024 reads, tile numbers from an atomic counter):
//       per read      filter (P1), contig/start/end, T>C bit mask over the concatenated alignment blocks (P3) --
//                     bit-parallel on 2-bit packed words for the PAR-CLIP shape (uniform length, one M op),
//                     a literal CIGAR walk otherwise
//       look-back #1  running max of (contig, end) over all earlier reads = the cluster end the Java loop holds
//                     (tempClusterEnd) -> boundary flag (clusterEnd - start) < 5 or contig changed (P2)
//       look-back #2  segmented reduction keyed by the flags: cluster number, event offset and the running
//                     aggregate of the open cluster (reads, T>C count, end, 51-bit mask, strand state P6)
//       output        the read that opens cluster c+1 writes the finished sums of cluster c (one writer per field,
//                     no atomics, no zero-initialised accumulators); T>C events (position, insertion key) are
//                     appended at the scanned offset, i.e. in read order
//   pl_site_kernel   one warp per cluster: mutationMap keys = distinct event positions, their counts, the
//                     first-insertion key and baseCoveredMap at those keys (difference array over the cluster's
//                     read intervals), in windows of 128 positions of shared memory; site slots come from a third
//                     look-back so the site array is compact and ordered by (cluster, position)
//
// The flush-time logic (SNP filter, anchor site, text rows: :178-344) stays on the host side of the boundary.
#include <algorithm>
#include <cstring>

#include "device_common.cuh"
#include "lookback.cuh"

void timer_begin(ps_ctx* ctx, cudaStream_t st);
void timer_end(ps_ctx* ctx, cudaStream_t st);
int stage_batch(ps_ctx* ctx, const ps_read_batch* hb, bool with_qual, StagedBatch** out);

namespace {

constexpr int PL_THREADS = 256;
constexpr int PL_SITE_CLUSTERS = 64;     // clusters per block of pl_site_kernel
constexpr int PL_WINDOW = 128;           // positions per shared-memory window

struct PlState {                  // device-side run state (one per call)
  unsigned long long fault;       // min over (ordinal << 8 | code); ~0 = none
  unsigned long long skipped;     // skippedDueIndel (:156)
  unsigned long long dstr;        // doubleStranded (:496)
  unsigned long long n_ev;        // T>C events
  unsigned long long n_sites;     // distinct (cluster, position)
  unsigned int n_flags;           // clusters opened
  unsigned int unsorted;
  unsigned int tile_ctr_scan;
  unsigned int tile_ctr_site;
};

// running aggregate of the cluster that is open after a prefix of the reads
struct Seg {
  uint32_t nflags;      // boundary flags in the prefix
  uint32_t reads;       // numReadsPerCluster
  uint32_t t2c;         // numT2CMutationPerCluster
  uint32_t minus;       // minus-strand reads of the cluster
  int32_t end;          // tempClusterEnd
  uint32_t first_rev;   // tempIsReverse (strand of the first read)
  unsigned long long nev;    // events in the prefix
  unsigned long long mask;   // alleleFrequencyPositionsTemp
};
__device__ __forceinline__ Seg seg_identity() {
  Seg s;
  s.nflags = 0; s.reads = 0; s.t2c = 0; s.minus = 0; s.end = INT32_MIN; s.first_rev = 0; s.nev = 0; s.mask = 0;
  return s;
}
struct SegOp {
  __device__ __forceinline__ Seg operator()(const Seg& a, const Seg& b) const {
    Seg r;
    r.nflags = a.nflags + b.nflags;
    r.nev = a.nev + b.nev;
    if (b.nflags) {
      r.reads = b.reads; r.t2c = b.t2c; r.minus = b.minus; r.end = b.end; r.first_rev = b.first_rev; r.mask = b.mask;
    } else {
      r.reads = a.reads + b.reads; r.t2c = a.t2c + b.t2c; r.minus = a.minus + b.minus;
      r.end = a.end > b.end ? a.end : b.end;
      r.first_rev = a.reads ? a.first_rev : b.first_rev;
      r.mask = a.mask | b.mask;
    }
    return r;
  }
};
struct MaxOp {
  __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const {
    return a > b ? a : b;
  }
};
struct SumOp {
  __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const {
    return a + b;
  }
};

struct PlItem {          // what the pileup needs from one read
  unsigned long long key;    // (contig+1) << 32 | end; 0 = record not kept
  unsigned long long mask;   // T>C by index i over the strand-oriented concatenated blocks
  int32_t start, end, lo, hi;
  bool rev;
};

struct ScanParams {
  DeviceBatch b;
  DeviceRef ref;
  PlState* st;
  LbDesc<unsigned long long>* d_max;
  LbDesc<Seg>* d_seg;
  unsigned int epoch;
  uint32_t n_tiles;
  unsigned long long carry_key;
  uint32_t first_id;
  // outputs
  int2* iv;                 // [n] checkPosition interval of every read (empty for records not kept)
  ps_cluster* cl;           // [cap_cl] slot 0 = reads continuing the carry-in cluster
  uint32_t* cl_first;       // [cap_cl+1] first read of each slot
  uint32_t* cl_ev;          // [cap_cl+1] first event of each slot
  int32_t* ev_pos;          // [cap_ev]
  unsigned long long* ev_key;   // [cap_ev] (read ordinal << 6) | i
  uint64_t cap_cl, cap_ev;
};

// ---------------------------------------------------------------------------------------------------------
// Literal per-read routine (any CIGAR, any flag): PileupClusters.java:146-158, :585-673
// ---------------------------------------------------------------------------------------------------------
__device__ __noinline__ void pl_decode_generic(const ScanParams& P, uint64_t r, uint32_t meta, ReadOffsets off, PlItem& it) {
  const uint32_t flags = PS_META_FLAGS(meta), L = PS_META_LEN(meta), ncig = PS_META_NCIGAR(meta);
  const uint32_t* cig = P.b.cigar + off.cigar;
  it.key = 0; it.mask = 0; it.start = 0; it.end = 0; it.lo = 1; it.hi = 0; it.rev = false;
  if (flags & PS_RF_UNMAPPED) return;                                        // :146
  uint32_t R = 0, alen = 0;
  bool hasI = false, hasD = false, hasN = false;
  for (uint32_t e = 0; e < ncig; ++e) {
    const uint32_t c = __ldg(cig + e), op = c & 15u;
    hasI |= op == 1u; hasD |= op == 2u; hasN |= op == 3u;
    if (op_consumes_ref(op)) R += c >> 4;
    if (op_is_match(op)) alen += c >> 4;
  }
  if ((hasI || hasD) && hasN) {                                              // :152-157
    atomicAdd(&P.st->skipped, 1ull);
    return;
  }
  if (flags & PS_RF_POS_ZERO) { raise_fault(&P.st->fault, r, PS_THROW_REF_RANGE); return; }
  const uint64_t g0 = __ldg(P.b.ref_start + r);
  const uint32_t contig = g0 < P.ref.n_bases ? contig_of(P.ref, g0) : P.ref.n_contigs - 1;
  const uint64_t c_lo = __ldg(P.ref.contig_off + contig), c_hi = __ldg(P.ref.contig_off + contig + 1);
  const int32_t start = (int32_t)(g0 - c_lo) + 1;
  const int32_t end = start + (int32_t)R - 1;
  const bool rev = flags & PS_RF_REVERSE;
  const bool has_inv = flags & PS_RF_HAS_INVALID;
  const uint8_t* rb = P.b.bases2 + off.base;
  // alignment blocks (SAMUtils.getAlignmentBlocks): S,I advance the read; D,N the reference; H,P nothing.
  // The reference slices every block (read bases, then FASTA) before it looks at a single base (:593-604).
  int64_t rdp = 0, rfp = 0;
  for (uint32_t e = 0; e < ncig; ++e) {
    const uint32_t c = __ldg(cig + e), op = c & 15u;
    const int64_t n = c >> 4;
    if (op == 4u || op == 1u) rdp += n;
    else if (op == 2u || op == 3u) rfp += n;
    else if (op_is_match(op)) {
      if (rdp + n > (int64_t)L) { raise_fault(&P.st->fault, r, PS_THROW_BLOCK_RANGE); return; }
      if ((flags & PS_RF_REF_RANGE) || g0 + (uint64_t)(rfp + n) > c_hi || g0 >= P.ref.n_bases) {
        raise_fault(&P.st->fault, r, PS_THROW_REF_RANGE); return;
      }
      rdp += n; rfp += n;
    }
  }
  rdp = 0; rfp = 0;
  uint32_t j = 0;             // index over the concatenated blocks (forward orientation)
  unsigned long long mask = 0;
  for (uint32_t e = 0; e < ncig; ++e) {
    const uint32_t c = __ldg(cig + e), op = c & 15u;
    const int64_t n = c >> 4;
    if (op == 4u || op == 1u) rdp += n;
    else if (op == 2u || op == 3u) rfp += n;
    else if (op_is_match(op)) {
      for (int64_t z = 0; z < n; ++z, ++j) {
        const uint64_t g = g0 + (uint64_t)(rfp + z);
        const uint32_t p = (uint32_t)(rdp + z);
        if (ref_invalid_at(P.ref, g)) continue;
        const uint32_t a = ref_code_at(P.ref, g), bb = read_code_at(rb, p);
        // minus strand: both arrays are reverse-complemented, so ref T & read C there is ref A & read G here
        const bool hit = rev ? (a == 0u && bb == 2u) : (a == 3u && bb == 1u);
        if (!hit) continue;
        if (has_inv && read_pos_invalid(P.b, r / PS_TILE_READS, (uint32_t)(r % PS_TILE_READS), p)) continue;
        const uint32_t i = rev ? alen - 1 - j : j;
        if (i >= 51u) { raise_fault(&P.st->fault, r, PS_THROW_MASK51); return; }   // boolean[51] (:654)
        mask |= 1ull << i;
      }
      rdp += n; rfp += n;
    }
  }
  it.key = ((unsigned long long)(contig + 1) << 32) | (uint32_t)end;
  it.mask = mask;
  it.start = start; it.end = end; it.rev = rev;
  if (rev) { it.hi = end; it.lo = end - (int32_t)alen + 1; }
  else { it.lo = start; it.hi = start + (int32_t)alen - 1; }
}

// ---------------------------------------------------------------------------------------------------------
// PAR-CLIP shape: uniform length L <= 64, one M/=/X op of length L, no N call in the read.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t compress_even(uint32_t c) {   // even bits of c -> low 16 bits
  c = (c | (c >> 1)) & 0x33333333u;
  c = (c | (c >> 2)) & 0x0F0F0F0Fu;
  c = (c | (c >> 4)) & 0x00FF00FFu;
  c = (c | (c >> 8)) & 0x0000FFFFu;
  return c;
}

template <int NW>
__device__ __forceinline__ unsigned long long t2c_mask_fast(const DeviceRef& ref, uint32_t g0, uint32_t L, bool rev,
                                                            const uint32_t* __restrict__ brow_w, uint32_t bshift) {
  const uint32_t wi = g0 >> 4, sh = (g0 & 15u) * 2u;
  uint32_t w[NW + 1], bw[NW + 1];
#pragma unroll
  for (int k = 0; k <= NW; ++k) { w[k] = __ldg(ref.seq2 + wi + k); bw[k] = brow_w[k]; }
  const uint32_t ii = g0 >> 5, s1 = g0 & 31u;
  const uint32_t i0 = __ldg(ref.inv + ii), i1 = __ldg(ref.inv + ii + 1), i2 = NW > 2 ? __ldg(ref.inv + ii + 2) : 0u;
  unsigned long long m = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    const uint32_t rf = __funnelshift_r(w[k], w[k + 1], sh);
    const uint32_t rd = __funnelshift_r(bw[k], bw[k + 1], bshift);
    // forward: ref T (11) and read C (01);  reverse strand: ref A (00) and read G (10)
    const uint32_t h = rev ? (~(rf | (rf >> 1)) & (rd >> 1) & ~rd) : (rf & (rf >> 1) & rd & ~(rd >> 1));
    m |= (unsigned long long)compress_even(h & 0x55555555u) << (16 * k);
  }
  const unsigned long long inv = (unsigned long long)__funnelshift_r(i0, i1, s1) |
                                 ((unsigned long long)(NW > 2 ? __funnelshift_r(i1, i2, s1) : 0u) << 32);
  m &= ~inv;
  m &= L >= 64 ? ~0ull : ((1ull << L) - 1ull);
  if (rev) m = __brevll(m) >> (64u - L);      // i = alignedLength - 1 - j
  return m;
}

// block-wide exclusive scans over one value per thread (PL_THREADS threads); also return the block total
template <typename T, typename Op>
__device__ __forceinline__ T block_exclusive(const T& v, Op op, const T& identity, T* warp_tot /* [8] smem */, T& total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const T y = lb_shfl_up(x, d);
    if (lane >= (uint32_t)d) x = op(y, x);
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  T prefix = identity;
  T tot = identity;
#pragma unroll
  for (uint32_t w = 0; w < PL_THREADS / 32; ++w) {
    const T t = warp_tot[w];
    if (w < warp) prefix = op(prefix, t);
    tot = op(tot, t);
  }
  total = tot;
  // exclusive value of this thread: prefix of earlier warps, then the lanes before it
  T up = lb_shfl_up(x, 1);
  if (lane == 0) up = identity;
  __syncthreads();
  return op(prefix, up);
}

template <int ITEMS, int NW>   // NW > 0: PAR-CLIP fast decode; NW == 0: generic batch (ITEMS == 1)
__global__ void __launch_bounds__(PL_THREADS) pl_scan_kernel(const __grid_constant__ ScanParams P) {
  constexpr int TILE = PL_THREADS * ITEMS;
  __shared__ unsigned long long s_wmax[PL_THREADS / 32];
  __shared__ Seg s_wseg[PL_THREADS / 32];
  __shared__ uint64_t s_scan[8];
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_pmax;
  __shared__ Seg s_pseg;
  __shared__ __align__(16) uint32_t s_bases[NW > 0 ? (TILE * NW + 16) : 4];

  if (threadIdx.x == 0) s_tile = atomicAdd(&P.st->tile_ctr_scan, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t n = P.b.n_reads;
  const uint64_t r0 = (uint64_t)tile * TILE + (uint64_t)threadIdx.x * ITEMS;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  PlItem it[ITEMS];
  if constexpr (NW > 0) {
    // ---- stage the tile's packed bases (contiguous, 16-byte aligned) ----
    const uint32_t L = P.b.uniform_len, bpr = (L + 3) >> 2;
    const uint64_t t0 = (uint64_t)tile * TILE;
    const uint32_t n_here = (uint32_t)min((uint64_t)TILE, n - t0);
    const uint32_t nbytes = n_here * bpr;
    const uint4* src = reinterpret_cast<const uint4*>(P.b.bases2 + t0 * bpr);
    uint4* dst = reinterpret_cast<uint4*>(s_bases);
    for (uint32_t k = threadIdx.x; k < (nbytes + 15) / 16; k += PL_THREADS) dst[k] = __ldg(src + k);   // batch is padded
    uint32_t metas[ITEMS], starts[ITEMS], cigs[ITEMS];
    static_assert(ITEMS == 4, "the fast decode takes 4 reads per thread");
    if (r0 + 4 <= n) {
      const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(P.b.meta + r0));
      const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(P.b.ref_start + r0));
      const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(P.b.cigar + r0));
      metas[0] = m4.x; metas[1] = m4.y; metas[2] = m4.z; metas[3] = m4.w;
      starts[0] = s4.x; starts[1] = s4.y; starts[2] = s4.z; starts[3] = s4.w;
      cigs[0] = c4.x; cigs[1] = c4.y; cigs[2] = c4.z; cigs[3] = c4.w;
    } else {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const bool in = r0 + j < n;
        metas[j] = in ? __ldg(P.b.meta + r0 + j) : PS_MAKE_META(0, 0, PS_RF_UNMAPPED);
        starts[j] = in ? __ldg(P.b.ref_start + r0 + j) : 0u;
        cigs[j] = in ? __ldg(P.b.cigar + r0 + j) : 0u;
      }
    }
    // contig bounds of the thread's first read, reused while the reads stay inside
    uint64_t c_lo = 1, c_hi = 0;
    uint32_t contig = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const uint32_t meta = metas[j], flags = PS_META_FLAGS(meta), g0 = starts[j], cg = cigs[j];
      bool fast = (flags & ~PS_RF_REVERSE) == 0 && op_is_match(cg & 15u) && (cg >> 4) == L && PS_META_LEN(meta) == L;
      if (fast && !((uint64_t)g0 >= c_lo && (uint64_t)g0 < c_hi)) {
        if ((uint64_t)g0 < P.ref.n_bases) {
          contig = contig_of(P.ref, g0);
          c_lo = __ldg(P.ref.contig_off + contig);
          c_hi = __ldg(P.ref.contig_off + contig + 1);
        } else fast = false;
      }
      fast = fast && (uint64_t)g0 + L <= c_hi;
      if (fast) {
        const bool rev = (flags & PS_RF_REVERSE) != 0;
        const uint32_t q = threadIdx.x * ITEMS + j, boff = q * bpr;
        unsigned long long m = t2c_mask_fast<NW>(P.ref, g0, L, rev, s_bases + (boff >> 2), (boff & 3u) * 8u);
        const int32_t start = (int32_t)((uint64_t)g0 - c_lo) + 1, end = start + (int32_t)L - 1;
        if (L > 51u && (m >> 51)) { raise_fault(&P.st->fault, r0 + j, PS_THROW_MASK51); m = 0; it[j].key = 0; it[j].lo = 1; it[j].hi = 0; }
        else { it[j].key = ((unsigned long long)(contig + 1) << 32) | (uint32_t)end; it[j].lo = start; it[j].hi = end; }
        it[j].mask = m; it[j].start = start; it[j].end = end; it[j].rev = rev;
      } else if (r0 + j < n) {
        ReadOffsets off;
        off.base = (r0 + j) * (uint64_t)bpr; off.qual = (r0 + j) * (uint64_t)L; off.cigar = r0 + j;
        PlItem tmp;                       // the out-of-line routine gets an addressable copy; it[] stays in registers
        pl_decode_generic(P, r0 + j, meta, off, tmp);
        it[j] = tmp;
      } else {
        it[j].key = 0; it[j].mask = 0; it[j].start = 0; it[j].end = 0; it[j].lo = 1; it[j].hi = 0; it[j].rev = false;
      }
    }
  } else {
    const bool in_range = r0 < n;
    const uint32_t meta = in_range ? __ldg(P.b.meta + r0) : 0;
    const ReadOffsets off = read_offsets(P.b, tile, r0, meta, in_range, s_scan);
    if (in_range) {
      PlItem tmp;
      pl_decode_generic(P, r0, meta, off, tmp);
      it[0] = tmp;
    } else { it[0].key = 0; it[0].mask = 0; it[0].start = 0; it[0].end = 0; it[0].lo = 1; it[0].hi = 0; it[0].rev = false; }
  }
  // checkPosition intervals (baseCoveredMap support) for the site kernel
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
    if (r0 + j < n) P.iv[r0 + j] = make_int2(it[j].lo, it[j].hi);

  // ---- look-back #1: running max of (contig, end) -----------------------------------------------------------
  unsigned long long tmax = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) tmax = it[j].key > tmax ? it[j].key : tmax;
  unsigned long long block_max;
  const unsigned long long ex_max = block_exclusive<unsigned long long>(tmax, MaxOp(), 0ull, s_wmax, block_max);
  if (warp == 0) {
    if (tile == 0) {
      if (lane == 0) { lb_publish(&P.d_max[0], block_max, 2u, P.epoch); s_pmax = 0; }
    } else {
      if (lane == 0) lb_publish(&P.d_max[tile], block_max, 1u, P.epoch);
      const unsigned long long pre = lb_exclusive_prefix(P.d_max, (int)tile, P.epoch, MaxOp(), 0ull);
      if (lane == 0) {
        lb_publish(&P.d_max[tile], pre > block_max ? pre : block_max, 2u, P.epoch);
        s_pmax = pre;
      }
    }
  }
  __syncthreads();
  unsigned long long E = s_pmax > ex_max ? s_pmax : ex_max;
  E = E > P.carry_key ? E : P.carry_key;

  // ---- boundary flags (:175-176) and this thread's segment aggregate --------------------------------------------
  bool flag[ITEMS];
  Seg tseg = seg_identity();
  bool unsorted = false;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    flag[j] = false;
    const unsigned long long my = it[j].key;
    if (my) {
      const uint32_t pc = (uint32_t)(E >> 32), mc = (uint32_t)(my >> 32);
      if (E == 0) flag[j] = true;                       // tempClusterEnd = 0, tempClusterChr = "" (:118-120)
      else if (pc != mc) flag[j] = true;
      else flag[j] = ((int64_t)(int32_t)(uint32_t)E - (int64_t)it[j].start) < 5;
      unsorted |= pc > mc;                              // contig order went backwards: not coordinate sorted
      E = my > E ? my : E;
      Seg e;
      e.nflags = flag[j] ? 1u : 0u; e.reads = 1; e.t2c = __popcll(it[j].mask); e.minus = it[j].rev ? 1u : 0u;
      e.end = it[j].end; e.first_rev = it[j].rev ? 1u : 0u; e.nev = e.t2c; e.mask = it[j].mask;
      tseg = SegOp()(tseg, e);
    }
  }
  if (unsorted) P.st->unsorted = 1u;

  // ---- look-back #2: segmented aggregate ------------------------------------------------------------------------
  Seg block_seg;
  const Seg ex_seg = block_exclusive<Seg>(tseg, SegOp(), seg_identity(), s_wseg, block_seg);
  if (warp == 0) {
    if (tile == 0) {
      if (lane == 0) { lb_publish(&P.d_seg[0], block_seg, 2u, P.epoch); s_pseg = seg_identity(); }
    } else {
      if (lane == 0) lb_publish(&P.d_seg[tile], block_seg, 1u, P.epoch);
      const Seg pre = lb_exclusive_prefix(P.d_seg, (int)tile, P.epoch, SegOp(), seg_identity());
      if (lane == 0) {
        lb_publish(&P.d_seg[tile], SegOp()(pre, block_seg), 2u, P.epoch);
        s_pseg = pre;
      }
    }
  }
  __syncthreads();
  Seg run = SegOp()(s_pseg, ex_seg);   // state left by every read before this thread's first

  // ---- records: the opener of a cluster closes the previous one ------------------------------------------------
  unsigned long long dstr = 0;
  auto close_slot = [&](const Seg& s) {
    const uint32_t slot = s.nflags;
    const uint32_t maf = slot ? s.minus - s.first_rev : s.minus;
    if (slot < P.cap_cl) {
      ps_cluster* c = P.cl + slot;
      c->end = s.reads ? s.end : 0;
      c->num_reads = s.reads;
      c->num_t2c = s.t2c;
      c->minus_after_first = maf;
      c->first_reverse = (uint8_t)s.first_rev;
      // StrandOrientation state at flush (P6): minus-first -> "-"; plus-first -> "+/-" once a minus member came
      c->combined_strand = s.first_rev ? 1 : (maf ? 2 : 0);
      c->reserved = 0;
      c->mask51 = s.mask;
    }
    if (slot && !s.first_rev) dstr += maf;              // doubleStranded++ (:494-498), incl. the never-flushed last cluster
  };
  if (tile == 0 && threadIdx.x == 0) { P.cl_first[0] = 0; P.cl_ev[0] = 0; }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if (!it[j].key) continue;
    const uint64_t r = r0 + j;
    const uint32_t t2c = __popcll(it[j].mask);
    if (flag[j]) {
      close_slot(run);
      const uint32_t slot = run.nflags + 1;
      if (slot < P.cap_cl) {
        ps_cluster* c = P.cl + slot;
        c->first_read = r;
        c->running_id = P.first_id + slot;               // runningID++ then "cl_<id>_<chr>" (:355): first cluster is cl_2
        c->contig = (uint32_t)(it[j].key >> 32) - 1;
        c->start = it[j].start;
        P.cl_first[slot] = (uint32_t)r;
        P.cl_ev[slot] = (uint32_t)run.nev;
      }
      run.nflags = slot; run.reads = 1; run.t2c = t2c; run.minus = it[j].rev ? 1u : 0u; run.end = it[j].end;
      run.first_rev = it[j].rev ? 1u : 0u; run.mask = it[j].mask;
    } else {
      if (run.nflags == 0 && run.reads == 0) {           // first read continuing the carry-in cluster (halo merge)
        ps_cluster* c = P.cl;
        c->first_read = r; c->running_id = 0; c->contig = (uint32_t)(it[j].key >> 32) - 1; c->start = it[j].start;
        run.first_rev = it[j].rev ? 1u : 0u;
      }
      run.reads += 1; run.t2c += t2c; run.minus += it[j].rev ? 1u : 0u;
      run.end = it[j].end > run.end ? it[j].end : run.end;
      run.mask |= it[j].mask;
    }
    unsigned long long m = it[j].mask;
    unsigned long long o = run.nev;
    while (m) {
      const int i = __ffsll((long long)m) - 1;
      m &= m - 1;
      if (o < P.cap_ev) {
        P.ev_pos[o] = it[j].rev ? it[j].end - i : it[j].start + i;      // checkPosition (:638-643)
        P.ev_key[o] = (r << 6) | (unsigned)i;
      }
      ++o;
    }
    run.nev = o;
  }
  if (tile == P.n_tiles - 1 && threadIdx.x == PL_THREADS - 1) {     // state after the last read: the open cluster
    close_slot(run);
    const uint64_t n_slots = (uint64_t)run.nflags + 1;
    if (n_slots <= P.cap_cl) { P.cl_first[n_slots] = (uint32_t)n; P.cl_ev[n_slots] = (uint32_t)run.nev; }
    P.st->n_flags = run.nflags;
    P.st->n_ev = run.nev;
  }
  // warp-reduce doubleStranded
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) dstr += __shfl_down_sync(0xFFFFFFFFu, dstr, d);
  if (lane == 0 && dstr) atomicAdd(&P.st->dstr, dstr);
}

// ---------------------------------------------------------------------------------------------------------
// Sites
// ---------------------------------------------------------------------------------------------------------
struct SiteParams {
  PlState* st;
  LbDesc<unsigned long long>* d_cnt;
  unsigned int epoch;
  uint32_t n_slots;
  uint32_t n_tiles;
  const int2* iv;
  ps_cluster* cl;
  const uint32_t* cl_first;
  const uint32_t* cl_ev;
  const int32_t* ev_pos;
  const unsigned long long* ev_key;
  ps_site* sites;
};

struct WarpTables {     // one window of PL_WINDOW positions
  int32_t diff[PL_WINDOW];       // +1 at lo, -1 after hi  -> prefix sum = baseCoveredMap
  uint32_t cnt[PL_WINDOW];       // mutationMap
  uint32_t first[PL_WINDOW];     // first event (index inside the cluster) = first insertion
};

// distinct event positions of one cluster (whole warp)
__device__ __forceinline__ uint32_t site_count(const SiteParams& P, uint32_t e0, uint32_t e1, WarpTables& T) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t ne = e1 - e0;
  if (ne == 0) return 0;
  if (ne <= 32) {
    const bool have = lane < ne;
    const int32_t pos = have ? __ldg(P.ev_pos + e0 + lane) : INT32_MIN + (int32_t)lane;   // dummies are all different
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, pos);
    const bool leader = have && (uint32_t)(__ffs((int)peers) - 1) == lane;
    return __popc(__ballot_sync(0xFFFFFFFFu, leader));
  }
  int32_t mn = INT32_MAX, mx = INT32_MIN;
  for (uint32_t e = e0 + lane; e < e1; e += 32) { const int32_t p = __ldg(P.ev_pos + e); mn = min(mn, p); mx = max(mx, p); }
  mn = __reduce_min_sync(0xFFFFFFFFu, mn);
  mx = __reduce_max_sync(0xFFFFFFFFu, mx);
  uint32_t total = 0;
  for (int64_t w0 = mn; w0 <= mx; w0 += PL_WINDOW) {
    for (uint32_t k = lane; k < PL_WINDOW; k += 32) T.cnt[k] = 0;
    __syncwarp();
    for (uint32_t e = e0 + lane; e < e1; e += 32) {
      const int64_t d = (int64_t)__ldg(P.ev_pos + e) - w0;
      if (d >= 0 && d < PL_WINDOW) T.cnt[d] = 1;
    }
    __syncwarp();
    for (uint32_t k = lane; k < PL_WINDOW; k += 32) total += __popc(__ballot_sync(0xFFFFFFFFu, T.cnt[k] != 0));
    __syncwarp();
  }
  return total;
}

// full tables of one cluster, sites written in position order from `out`
__device__ __forceinline__ void site_emit(const SiteParams& P, uint32_t r_lo, uint32_t r_hi, uint32_t e0, uint32_t e1,
                                          WarpTables& T, ps_site* out) {
  const uint32_t lane = threadIdx.x & 31;
  int32_t mn = INT32_MAX, mx = INT32_MIN;
  for (uint32_t e = e0 + lane; e < e1; e += 32) { const int32_t p = __ldg(P.ev_pos + e); mn = min(mn, p); mx = max(mx, p); }
  mn = __reduce_min_sync(0xFFFFFFFFu, mn);
  mx = __reduce_max_sync(0xFFFFFFFFu, mx);
  uint32_t written = 0;
  for (int64_t w0 = mn; w0 <= mx; w0 += PL_WINDOW) {
    const int64_t w1 = w0 + PL_WINDOW - 1;
    for (uint32_t k = lane; k < PL_WINDOW; k += 32) { T.diff[k] = 0; T.cnt[k] = 0; T.first[k] = 0xFFFFFFFFu; }
    __syncwarp();
    for (uint32_t e = e0 + lane; e < e1; e += 32) {
      const int64_t d = (int64_t)__ldg(P.ev_pos + e) - w0;
      if (d >= 0 && d < PL_WINDOW) { atomicAdd(&T.cnt[d], 1u); atomicMin(&T.first[d], e - e0); }
    }
    for (uint32_t r = r_lo + lane; r < r_hi; r += 32) {
      const int2 v = __ldg(P.iv + r);
      if (v.x > v.y || (int64_t)v.y < w0 || (int64_t)v.x > w1) continue;
      atomicAdd(&T.diff[(int64_t)v.x > w0 ? (int64_t)v.x - w0 : 0], 1);
      if ((int64_t)v.y < w1) atomicAdd(&T.diff[(int64_t)v.y + 1 - w0], -1);
    }
    __syncwarp();
    int32_t carry = 0;
    for (uint32_t k0 = 0; k0 < PL_WINDOW; k0 += 32) {
      int32_t c = T.diff[k0 + lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int32_t y = __shfl_up_sync(0xFFFFFFFFu, c, d);
        if (lane >= (uint32_t)d) c += y;
      }
      c += carry;
      carry = __shfl_sync(0xFFFFFFFFu, c, 31);
      const uint32_t n = T.cnt[k0 + lane];
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, n != 0);
      if (n) {
        ps_site s;
        s.pos = (int32_t)(w0 + k0 + lane);
        s.t2c = n;
        s.cov = (uint32_t)c;
        s.reserved = 0;
        s.order_key = __ldg(P.ev_key + e0 + T.first[k0 + lane]);
        out[written + __popc(bal & ((1u << lane) - 1u))] = s;
      }
      written += __popc(bal);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(PL_THREADS) pl_site_kernel(const __grid_constant__ SiteParams P) {
  __shared__ WarpTables s_tab[PL_THREADS / 32];
  __shared__ uint32_t s_cnt[PL_SITE_CLUSTERS], s_off[PL_SITE_CLUSTERS];
  __shared__ unsigned long long s_base;
  __shared__ unsigned int s_tile;
  if (threadIdx.x == 0) s_tile = atomicAdd(&P.st->tile_ctr_site, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t c0 = tile * PL_SITE_CLUSTERS;
  WarpTables& T = s_tab[warp];
  for (uint32_t k = warp; k < PL_SITE_CLUSTERS; k += PL_THREADS / 32) {
    const uint32_t c = c0 + k;
    uint32_t cnt = 0;
    if (c < P.n_slots) cnt = site_count(P, __ldg(P.cl_ev + c), __ldg(P.cl_ev + c + 1), T);
    if (lane == 0) s_cnt[k] = cnt;
  }
  __syncthreads();
  if (warp == 0) {
    // exclusive prefix of the 64 counts (2 per lane) and the tile's base from the look-back
    const uint32_t a = s_cnt[2 * lane], b = s_cnt[2 * lane + 1];
    uint32_t x = a + b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
      if (lane >= (uint32_t)d) x += y;
    }
    const unsigned long long total = __shfl_sync(0xFFFFFFFFu, x, 31);
    s_off[2 * lane] = x - a - b;
    s_off[2 * lane + 1] = x - b;
    unsigned long long pre = 0;
    if (tile == 0) {
      if (lane == 0) lb_publish(&P.d_cnt[0], total, 2u, P.epoch);
    } else {
      if (lane == 0) lb_publish(&P.d_cnt[tile], total, 1u, P.epoch);
      pre = lb_exclusive_prefix(P.d_cnt, (int)tile, P.epoch, SumOp(), 0ull);
      if (lane == 0) lb_publish(&P.d_cnt[tile], pre + total, 2u, P.epoch);
    }
    if (lane == 0) {
      s_base = pre;
      if (tile == P.n_tiles - 1) P.st->n_sites = pre + total;
    }
  }
  __syncthreads();
  const unsigned long long base = s_base;
  for (uint32_t k = warp; k < PL_SITE_CLUSTERS; k += PL_THREADS / 32) {
    const uint32_t c = c0 + k;
    if (c >= P.n_slots) break;
    const unsigned long long b = base + s_off[k];
    const uint32_t cnt = s_cnt[k];
    if (cnt) site_emit(P, __ldg(P.cl_first + c), __ldg(P.cl_first + c + 1), __ldg(P.cl_ev + c), __ldg(P.cl_ev + c + 1), T, P.sites + b);
    if (lane == 0) { P.cl[c].site_begin = b; P.cl[c].site_end = b + cnt; }
  }
}

__global__ void pl_init_state(PlState* st) {
  if (threadIdx.x == 0) {
    st->fault = PS_FAULT_NONE; st->skipped = 0; st->dstr = 0; st->n_ev = 0; st->n_sites = 0; st->n_flags = 0;
    st->unsorted = 0; st->tile_ctr_scan = 0; st->tile_ctr_site = 0;
  }
}

}  // namespace

struct ps_pileup {
  ps_ctx* ctx = nullptr;
  uint64_t generation = 0;           // ctx->pl_generation at creation: scratch (iv, cl_first) is valid while equal
  cudaStream_t stream = nullptr;
  // device results, owned by the handle (stream-ordered allocations)
  ps_cluster* d_cl = nullptr;        // slot 0 = head partial, 1..n_flags-1 closed, n_flags = open
  ps_site* d_sites = nullptr;
  uint64_t n_slots = 0, n_reads = 0;
  ps_cluster head{}, open{};
  bool has_head = false;
  std::vector<uint32_t> open_cov, head_cov;
  int32_t open_cov_pos0 = 0, head_cov_pos0 = 0;
  bool cov_done = false;
  ps_pileup_counters counters{};
  ps_fault fault{};
};

template <typename T>
static T* scratch(ps_ctx* ctx, int slot, size_t count, cudaError_t& err, bool zero_new = false) {
  if (err != cudaSuccess) return nullptr;
  const size_t bytes = count * sizeof(T) + 64;
  const bool grow = bytes > ctx->pl_scratch[slot].cap;
  err = ctx->pl_scratch[slot].reserve(bytes);
  if (err == cudaSuccess && grow && zero_new) err = cudaMemset(ctx->pl_scratch[slot].p, 0, ctx->pl_scratch[slot].cap);
  return static_cast<T*>(ctx->pl_scratch[slot].p);
}

static bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int run_pileup(ps_ctx* ctx, const DeviceBatch& b, const ps_pileup_opts* opts, cudaStream_t st, ps_pileup** out) {
  ps_pileup* H = new ps_pileup();
  *out = H;
  H->ctx = ctx;
  H->stream = st;
  const uint64_t n = b.n_reads;
  H->n_reads = n;
  H->counters.num_reads_processed = n;
  H->generation = ++ctx->pl_generation;
  if (n == 0) return PS_OK;
  if (n >= 0xFFFFFFFFull) return set_error(ctx, PS_ERR_UNSUPPORTED, "pileup batch of >= 2^32 reads");

  const uint32_t L = b.uniform_len;
  const bool fast = L >= 1 && L <= 64 && b.uniform_ncigar == 1 && aligned16p(b.meta) && aligned16p(b.ref_start) &&
                    aligned16p(b.cigar) && aligned16p(b.bases2);
  const uint32_t items = fast ? 4u : 1u;
  const uint32_t tile_reads = PL_THREADS * items;
  const uint32_t n_tiles = (uint32_t)((n + tile_reads - 1) / tile_reads);

  cudaError_t err = cudaSuccess;
  PlState* d_state = scratch<PlState>(ctx, 0, 1, err);
  int2* iv = scratch<int2>(ctx, 1, n, err);
  auto* d_max = scratch<LbDesc<unsigned long long>>(ctx, 2, n_tiles, err, true);
  auto* d_seg = scratch<LbDesc<Seg>>(ctx, 3, n_tiles, err, true);
  if (err != cudaSuccess) return cuda_fail(ctx, err, "pileup scratch");
  if (ctx->pl_cap_cl < 1024) ctx->pl_cap_cl = std::max<uint64_t>(1024, n / 8);
  if (ctx->pl_cap_ev < 1024) ctx->pl_cap_ev = std::max<uint64_t>(1024, n);

  unsigned long long carry_key = 0;
  if (opts && opts->carry_valid)
    carry_key = ((unsigned long long)(opts->carry_contig + 1) << 32) | (uint32_t)opts->carry_cluster_end;

  PlState hs{};
  timer_begin(ctx, st);
  for (int attempt = 0;; ++attempt) {
    const uint64_t cap_cl = std::min<uint64_t>(ctx->pl_cap_cl, n + 2), cap_ev = ctx->pl_cap_ev;
    uint32_t* cl_first = scratch<uint32_t>(ctx, 4, cap_cl + 2, err);
    uint32_t* cl_ev = scratch<uint32_t>(ctx, 5, cap_cl + 2, err);
    int32_t* ev_pos = scratch<int32_t>(ctx, 6, cap_ev, err);
    unsigned long long* ev_key = scratch<unsigned long long>(ctx, 7, cap_ev, err);
    if (err != cudaSuccess) return cuda_fail(ctx, err, "pileup scratch");
    if (H->d_cl) { cudaFreeAsync(H->d_cl, st); H->d_cl = nullptr; }
    PS_CUDA(ctx, cudaMallocAsync((void**)&H->d_cl, cap_cl * sizeof(ps_cluster), st));

    ScanParams P;
    P.b = b; P.ref = ctx->ref; P.st = d_state; P.d_max = d_max; P.d_seg = d_seg; P.epoch = ++ctx->pl_epoch;
    P.n_tiles = n_tiles; P.carry_key = carry_key; P.first_id = opts ? opts->first_running_id : 1;
    P.iv = iv; P.cl = H->d_cl; P.cl_first = cl_first; P.cl_ev = cl_ev; P.ev_pos = ev_pos; P.ev_key = ev_key;
    P.cap_cl = cap_cl; P.cap_ev = cap_ev;
    pl_init_state<<<1, 32, 0, st>>>(d_state);
    if (!fast) pl_scan_kernel<1, 0><<<n_tiles, PL_THREADS, 0, st>>>(P);
    else if (L <= 16) pl_scan_kernel<4, 1><<<n_tiles, PL_THREADS, 0, st>>>(P);
    else if (L <= 32) pl_scan_kernel<4, 2><<<n_tiles, PL_THREADS, 0, st>>>(P);
    else if (L <= 48) pl_scan_kernel<4, 3><<<n_tiles, PL_THREADS, 0, st>>>(P);
    else pl_scan_kernel<4, 4><<<n_tiles, PL_THREADS, 0, st>>>(P);
    ctx->launches += 2;
    PS_CUDA(ctx, cudaGetLastError());
    PS_CUDA(ctx, cudaMemcpyAsync(&hs, d_state, sizeof(PlState), cudaMemcpyDeviceToHost, st));
    PS_CUDA(ctx, cudaStreamSynchronize(st));
    const uint64_t need_cl = (uint64_t)hs.n_flags + 2, need_ev = hs.n_ev;
    if (need_cl <= cap_cl && need_ev <= cap_ev) break;
    if (attempt >= 1) { timer_end(ctx, st); return set_error(ctx, PS_ERR_CUDA, "pileup: capacity retry failed"); }
    // totals do not depend on the capacities (dropped writes only): size exactly and run once more
    ctx->pl_cap_cl = std::max(ctx->pl_cap_cl, need_cl + need_cl / 16);
    ctx->pl_cap_ev = std::max(ctx->pl_cap_ev, need_ev + need_ev / 16);
  }
  H->counters.skipped_due_indel = hs.skipped;
  if (hs.fault != PS_FAULT_NONE) {
    H->fault.code = (int32_t)(hs.fault & 0xFF);
    H->fault.read_ordinal = hs.fault >> 8;
    timer_end(ctx, st);
    return set_error(ctx, PS_ERR_REFERENCE_WOULD_THROW, "pileup: the JVM would die on a record of this batch");
  }
  if (hs.unsorted) { timer_end(ctx, st); return set_error(ctx, PS_ERR_UNSORTED, ps_strerror(PS_ERR_UNSORTED)); }

  const uint64_t n_slots = (uint64_t)hs.n_flags + 1;
  H->n_slots = n_slots;
  const uint32_t s_tiles = (uint32_t)((n_slots + PL_SITE_CLUSTERS - 1) / PL_SITE_CLUSTERS);
  auto* d_cnt = scratch<LbDesc<unsigned long long>>(ctx, 8, s_tiles, err, true);
  if (err != cudaSuccess) return cuda_fail(ctx, err, "pileup scratch");
  PS_CUDA(ctx, cudaMallocAsync((void**)&H->d_sites, (hs.n_ev + 1) * sizeof(ps_site), st));
  SiteParams Q;
  Q.st = d_state; Q.d_cnt = d_cnt; Q.epoch = ++ctx->pl_epoch; Q.n_slots = (uint32_t)n_slots; Q.n_tiles = s_tiles;
  Q.iv = iv; Q.cl = H->d_cl; Q.cl_first = (const uint32_t*)ctx->pl_scratch[4].p; Q.cl_ev = (const uint32_t*)ctx->pl_scratch[5].p;
  Q.ev_pos = (const int32_t*)ctx->pl_scratch[6].p; Q.ev_key = (const unsigned long long*)ctx->pl_scratch[7].p;
  Q.sites = H->d_sites;
  pl_site_kernel<<<s_tiles, PL_THREADS, 0, st>>>(Q);
  ctx->launches += 1;
  PS_CUDA(ctx, cudaGetLastError());
  timer_end(ctx, st);
  // summary: state, head partial (slot 0), open cluster (last slot)
  PS_CUDA(ctx, cudaMemcpyAsync(&hs, d_state, sizeof(PlState), cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaMemcpyAsync(&H->head, H->d_cl, sizeof(ps_cluster), cudaMemcpyDeviceToHost, st));
  if (hs.n_flags) PS_CUDA(ctx, cudaMemcpyAsync(&H->open, H->d_cl + hs.n_flags, sizeof(ps_cluster), cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  H->has_head = H->head.num_reads != 0;
  H->counters.has_open_cluster = hs.n_flags ? 1 : 0;
  H->counters.double_stranded = hs.dstr;
  H->counters.n_clusters = hs.n_flags ? hs.n_flags - 1 : 0;
  const uint64_t sites_before_open = hs.n_flags ? H->open.site_begin : hs.n_sites;
  H->counters.n_sites = sites_before_open - H->head.site_end;
  return PS_OK;
}

static DeviceBatch pl_view_of(const ps_read_batch* b) {
  DeviceBatch v;
  v.n_reads = b->n_reads; v.meta = b->meta; v.ref_start = b->ref_start; v.bases2 = b->bases2; v.qual = b->qual;
  v.cigar = b->cigar; v.tile_base_off = b->tile_base_off; v.tile_qual_off = b->tile_qual_off;
  v.tile_cigar_off = b->tile_cigar_off; v.tile_exc_off = b->tile_exc_off; v.exc = b->exc;
  v.uniform_len = b->uniform_len; v.uniform_ncigar = b->uniform_ncigar;
  return v;
}

static void free_handle(ps_pileup* h) {
  if (!h) return;
  if (h->ctx) cudaSetDevice(h->ctx->device);
  if (h->d_cl) cudaFreeAsync(h->d_cl, h->stream);
  if (h->d_sites) cudaFreeAsync(h->d_sites, h->stream);
  delete h;
}

// dense baseCoveredMap of the two boundary clusters, from the read intervals (needed only by the halo merge)
static int boundary_coverage(ps_pileup* h) {
  if (h->cov_done) return PS_OK;
  ps_ctx* ctx = h->ctx;
  if (!ctx || h->n_reads == 0) { h->cov_done = true; return PS_OK; }
  if (ctx->pl_generation != h->generation)
    return set_error(ctx, PS_ERR_STATE, "boundary coverage must be read before the next pileup call on this context");
  cudaSetDevice(ctx->device);
  const int2* iv = (const int2*)ctx->pl_scratch[1].p;
  const uint32_t* cl_first = (const uint32_t*)ctx->pl_scratch[4].p;
  auto dense = [&](uint64_t r0, uint64_t r1, std::vector<uint32_t>& cov, int32_t& pos0) -> int {
    if (r1 <= r0) return PS_OK;
    std::vector<int2> v(r1 - r0);
    PS_CUDA(ctx, cudaMemcpy(v.data(), iv + r0, (r1 - r0) * sizeof(int2), cudaMemcpyDeviceToHost));
    int32_t mn = INT32_MAX, mx = INT32_MIN;
    for (const int2& x : v)
      if (x.x <= x.y) { mn = std::min(mn, x.x); mx = std::max(mx, x.y); }
    if (mn > mx) return PS_OK;
    pos0 = mn;
    cov.assign((size_t)(mx - mn + 1), 0);
    for (const int2& x : v)
      for (int32_t p = x.x; p <= x.y; ++p) cov[p - mn]++;
    return PS_OK;
  };
  const uint64_t n_flags = h->n_slots - 1;
  if (h->has_head) {
    uint32_t r1 = (uint32_t)h->n_reads;
    if (n_flags) PS_CUDA(ctx, cudaMemcpy(&r1, cl_first + 1, 4, cudaMemcpyDeviceToHost));
    int rc = dense(0, r1, h->head_cov, h->head_cov_pos0);
    if (rc) return rc;
  }
  if (h->counters.has_open_cluster) {
    int rc = dense(h->open.first_read, h->n_reads, h->open_cov, h->open_cov_pos0);
    if (rc) return rc;
  }
  h->cov_done = true;
  return PS_OK;
}

extern "C" {

int ps_pileup_batch_device(ps_ctx* ctx, const ps_read_batch* b, const ps_pileup_opts* opts, void* stream,
                           ps_pileup** out) {
  if (!ctx || !b || !out) return PS_ERR_INVALID_ARG;
  *out = nullptr;
  if (!ctx->ref_loaded) return set_error(ctx, PS_ERR_STATE, "no reference loaded");
  cudaSetDevice(ctx->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  int rc = run_pileup(ctx, pl_view_of(b), opts, st, out);
  if (rc != PS_OK && rc != PS_ERR_REFERENCE_WOULD_THROW) { free_handle(*out); *out = nullptr; }
  return rc;
}

int ps_pileup_batch(ps_ctx* ctx, const ps_read_batch* hb, const ps_pileup_opts* opts, ps_pileup** out) {
  if (!ctx || !hb || !out) return PS_ERR_INVALID_ARG;
  *out = nullptr;
  if (!ctx->ref_loaded) return set_error(ctx, PS_ERR_STATE, "no reference loaded");
  cudaSetDevice(ctx->device);
  if (hb->n_reads == 0) { *out = new ps_pileup(); (*out)->ctx = ctx; return PS_OK; }
  StagedBatch* sb = nullptr;
  int rc = stage_batch(ctx, hb, /*with_qual=*/false, &sb);
  if (rc) return rc;
  rc = run_pileup(ctx, sb->view, opts, ctx->stream, out);
  if (rc != PS_OK && rc != PS_ERR_REFERENCE_WOULD_THROW) { free_handle(*out); *out = nullptr; }
  return rc;
}

int ps_pileup_counters_get(const ps_pileup* h, ps_pileup_counters* out) {
  if (!h || !out) return PS_ERR_INVALID_ARG;
  *out = h->counters;
  return PS_OK;
}

int64_t ps_pileup_next(ps_pileup* h, uint64_t first, ps_cluster* clusters, uint64_t max_clusters, ps_site* sites,
                       uint64_t max_sites) {
  if (!h || !clusters || (!sites && max_sites)) return PS_ERR_INVALID_ARG;
  const uint64_t n_closed = h->counters.n_clusters;
  if (first >= n_closed || max_clusters == 0) return 0;
  ps_ctx* ctx = h->ctx;
  cudaSetDevice(ctx->device);
  uint64_t cnt = std::min<uint64_t>(max_clusters, n_closed - first);
  // closed clusters are slots 1 .. n_flags-1; their sites are contiguous and in slot order
  PS_CUDA(ctx, cudaMemcpyAsync(clusters, h->d_cl + 1 + first, cnt * sizeof(ps_cluster), cudaMemcpyDeviceToHost, h->stream));
  PS_CUDA(ctx, cudaStreamSynchronize(h->stream));
  const uint64_t sb = clusters[0].site_begin;
  uint64_t m = 0;
  while (m < cnt && clusters[m].site_end - sb <= max_sites) ++m;
  if (m == 0) return 0;
  const uint64_t ns = clusters[m - 1].site_end - sb;
  if (ns) {
    PS_CUDA(ctx, cudaMemcpyAsync(sites, h->d_sites + sb, ns * sizeof(ps_site), cudaMemcpyDeviceToHost, h->stream));
    PS_CUDA(ctx, cudaStreamSynchronize(h->stream));
  }
  for (uint64_t k = 0; k < m; ++k) { clusters[k].site_begin -= sb; clusters[k].site_end -= sb; }
  return (int64_t)m;
}

static int copy_boundary(ps_pileup* h, const ps_cluster& src, ps_cluster* cluster, ps_site* sites, uint64_t max_sites) {
  const uint64_t ns = src.site_end - src.site_begin;
  if (ns > max_sites) return PS_ERR_INVALID_ARG;
  *cluster = src;
  cluster->site_begin = 0;
  cluster->site_end = ns;
  if (ns) {
    if (!sites) return PS_ERR_INVALID_ARG;
    ps_ctx* ctx = h->ctx;
    cudaSetDevice(ctx->device);
    PS_CUDA(ctx, cudaMemcpyAsync(sites, h->d_sites + src.site_begin, ns * sizeof(ps_site), cudaMemcpyDeviceToHost, h->stream));
    PS_CUDA(ctx, cudaStreamSynchronize(h->stream));
  }
  return 1;
}

int ps_pileup_open_cluster(ps_pileup* h, ps_cluster* cluster, ps_site* sites, uint64_t max_sites) {
  if (!h || !cluster) return PS_ERR_INVALID_ARG;
  if (!h->counters.has_open_cluster) return 0;
  return copy_boundary(h, h->open, cluster, sites, max_sites);
}

int ps_pileup_head_partial(ps_pileup* h, ps_cluster* cluster, ps_site* sites, uint64_t max_sites) {
  if (!h || !cluster) return PS_ERR_INVALID_ARG;
  if (!h->has_head) return 0;
  return copy_boundary(h, h->head, cluster, sites, max_sites);
}

int64_t ps_pileup_boundary_coverage(ps_pileup* h, int which, int32_t* first_pos, uint32_t* cov, uint64_t max) {
  if (!h || !first_pos) return PS_ERR_INVALID_ARG;
  int rc = boundary_coverage(h);
  if (rc) return rc;
  const std::vector<uint32_t>& v = which ? h->open_cov : h->head_cov;
  *first_pos = which ? h->open_cov_pos0 : h->head_cov_pos0;
  if (!cov) return (int64_t)v.size();
  if (v.size() > max) return PS_ERR_INVALID_ARG;
  if (!v.empty()) std::memcpy(cov, v.data(), v.size() * 4);
  return (int64_t)v.size();
}

int ps_pileup_fault(const ps_pileup* h, ps_fault* out) {
  if (!h || !out) return PS_ERR_INVALID_ARG;
  *out = h->fault;
  return PS_OK;
}

void ps_pileup_close(ps_pileup* h) { free_handle(h); }

}  // extern "C"
