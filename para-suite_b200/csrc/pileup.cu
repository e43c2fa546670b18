// K2-K4: T>C conversion pileup.  Replaces the loop PileupClusters.java:137-500 and
// calculateClusterInformation :585-673 (reference: /root/reference/src/src/utils/pileupclusters/).
//
// Pipeline over one coordinate-sorted SoA batch (all on one stream):
//   pl_read_kernel   per read: filter (P1), contig/start/end, T>C bit mask over the concatenated alignment
//                    blocks (P3), checkPosition range [lo,hi], scan key (contig+1)<<32|end
//   max-scan         E_k = running max of the keys  (cluster end seen so far on this contig)
//   pl_flag_kernel   boundary flag: (E_{k-1}.end - start_k) < 5 or contig changed (P2); sortedness check
//   sum-scan         cluster index per read; event offsets per read
//   pl_cluster_kernel per-cluster reductions (reads, T>C count, end, 51-bit mask, strand state P6) with
//                    warp-segmented aggregation before the global atomics; T>C events (cluster,pos)->order key
//   sort + pl_site_kernel  unique (cluster,pos) = mutationMap keys; count, first-insertion key, coverage
//                    (baseCoveredMap is only ever read at mutationMap keys: PileupClusters.java:215)
// The flush-time logic (SNP filter, anchor site, text rows: :178-344) stays on the host side of the boundary.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>

#include "device_common.cuh"

void timer_begin(ps_ctx* ctx, cudaStream_t st);
void timer_end(ps_ctx* ctx, cudaStream_t st);
int stage_batch(ps_ctx* ctx, const ps_read_batch* hb, StagedBatch** out);

struct ps_pileup {
  std::vector<ps_cluster> clusters;   // closed clusters
  std::vector<ps_site> sites;
  ps_cluster open_cluster{};
  std::vector<ps_site> open_sites;
  ps_cluster head_partial{};          // reads continuing the carry-in cluster (halo merge)
  std::vector<ps_site> head_sites;
  bool has_head = false;
  // baseCoveredMap of the two boundary clusters as dense arrays (needed by the halo merge: a site seen on one
  // side of a cut is also covered by reads on the other side)
  std::vector<uint32_t> open_cov, head_cov;
  int32_t open_cov_pos0 = 0, head_cov_pos0 = 0;
  ps_pileup_counters counters{};
  ps_fault fault{};
};

namespace {

struct PlRead {          // per-read scratch (SoA on the device)
  unsigned long long* key;      // scan key, 0 for reads that are not kept
  unsigned long long* t2c;      // T>C mask by index i over the (strand-oriented) concatenated blocks
  int32_t* start;               // 1-based
  int32_t* end;
  int32_t* lo;                  // checkPosition range
  int32_t* hi;
  uint32_t* flag;               // 1 = opens a cluster
  uint32_t* nev;                // T>C events of this read
};

struct PlParams {
  DeviceBatch b;
  DeviceRef ref;
  PlRead rd;
  unsigned long long* fault;
  unsigned long long* counters;   // [0] skipped_due_indel, [1] unsorted flag
  uint32_t n_tiles;
};

__global__ void __launch_bounds__(PS_BLOCK_THREADS) pl_read_kernel(const PlParams P) {
  __shared__ uint64_t s_scan[8];
  for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const uint64_t r = (uint64_t)tile * PS_TILE_READS + threadIdx.x;
    const bool in_range = r < P.b.n_reads;
    const uint32_t meta = in_range ? __ldg(P.b.meta + r) : 0;
    const ReadOffsets off = read_offsets(P.b, tile, r, meta, in_range, s_scan);
    if (!in_range) continue;
    const uint32_t flags = PS_META_FLAGS(meta), L = PS_META_LEN(meta), ncig = PS_META_NCIGAR(meta);
    const uint32_t* cig = P.b.cigar + off.cigar;
    unsigned long long key = 0, mask = 0;
    int32_t start = 0, end = 0, lo = 1, hi = 0;
    bool keep = !(flags & PS_RF_UNMAPPED);                                   // :146
    uint32_t R = 0, alen = 0;
    if (keep) {
      bool hasI = false, hasD = false, hasN = false;
      for (uint32_t e = 0; e < ncig; ++e) {
        const uint32_t c = __ldg(cig + e), op = c & 15u;
        hasI |= op == 1u; hasD |= op == 2u; hasN |= op == 3u;
        if (op_consumes_ref(op)) R += c >> 4;
        if (op_is_match(op)) alen += c >> 4;
      }
      if ((hasI || hasD) && hasN) {                                          // :152-157
        atomicAdd(&P.counters[0], 1ull);
        keep = false;
      }
    }
    if (keep && (flags & PS_RF_POS_ZERO)) { raise_fault(P.fault, r, PS_THROW_REF_RANGE); keep = false; }
    if (keep) {
      const uint64_t g0 = __ldg(P.b.ref_start + r);
      const uint32_t contig = g0 < P.ref.n_bases ? contig_of(P.ref, g0) : P.ref.n_contigs - 1;
      const uint64_t c_lo = __ldg(P.ref.contig_off + contig), c_hi = __ldg(P.ref.contig_off + contig + 1);
      start = (int32_t)(g0 - c_lo) + 1;
      end = start + (int32_t)R - 1;
      const bool rev = flags & PS_RF_REVERSE;
      const bool has_inv = flags & PS_RF_HAS_INVALID;
      const uint8_t* rb = P.b.bases2 + off.base;
      // alignment blocks (SAMUtils.getAlignmentBlocks): S,I advance the read; D,N the reference; H,P nothing
      int64_t rdp = 0, rfp = 0;   // read cursor, reference cursor relative to g0
      uint32_t j = 0;             // index over the concatenated blocks (forward orientation)
      bool dead = false;
      // the reference slices every block (read bases, then FASTA) before it looks at a single base (:593-604)
      for (uint32_t e = 0; e < ncig && !dead; ++e) {
        const uint32_t c = __ldg(cig + e), op = c & 15u;
        const int64_t n = c >> 4;
        if (op == 4u || op == 1u) rdp += n;
        else if (op == 2u || op == 3u) rfp += n;
        else if (op_is_match(op)) {
          if (rdp + n > (int64_t)L) { raise_fault(P.fault, r, PS_THROW_BLOCK_RANGE); dead = true; break; }
          if ((flags & PS_RF_REF_RANGE) || g0 + (uint64_t)(rfp + n) > c_hi || g0 >= P.ref.n_bases) {
            raise_fault(P.fault, r, PS_THROW_REF_RANGE); dead = true; break;
          }
          rdp += n; rfp += n;
        }
      }
      rdp = 0; rfp = 0;
      for (uint32_t e = 0; e < ncig && !dead; ++e) {
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
            if (has_inv && read_pos_invalid(P.b, tile, threadIdx.x, p)) continue;
            const uint32_t i = rev ? alen - 1 - j : j;
            if (i >= 51u) {   // mutationMapInRead[i] = true on boolean[51]  (:654)
              // the reference hits the lowest i first; report the read, the code is the same
              raise_fault(P.fault, r, PS_THROW_MASK51); dead = true; break;
            }
            mask |= 1ull << i;
          }
          rdp += n; rfp += n;
        }
      }
      if (dead) { keep = false; mask = 0; }
      else {
        if (rev) { hi = end; lo = end - (int32_t)alen + 1; }
        else { lo = start; hi = start + (int32_t)alen - 1; }
        key = ((unsigned long long)(contig + 1) << 32) | (uint32_t)end;
      }
    }
    P.rd.key[r] = key;
    P.rd.t2c[r] = mask;
    P.rd.start[r] = start;
    P.rd.end[r] = end;
    P.rd.lo[r] = lo;
    P.rd.hi[r] = hi;
    P.rd.nev[r] = __popcll(mask);
  }
}

struct MaxOp {
  __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const {
    return a > b ? a : b;
  }
};

// boundary flags from the inclusive running max E (E[k-1] = state left by all earlier records)
__global__ void pl_flag_kernel(uint64_t n, const unsigned long long* key, const unsigned long long* E,
                               const int32_t* start, uint32_t* flag, unsigned long long carry_key,
                               unsigned long long* counters) {
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long my = key[k];
    uint32_t f = 0;
    if (my) {
      const unsigned long long prev = k ? E[k - 1] : 0ull;
      const unsigned long long p = prev > carry_key ? prev : carry_key;
      const uint32_t pc = (uint32_t)(p >> 32), mc = (uint32_t)(my >> 32);
      if (p == 0) f = 1;                               // tempClusterEnd = 0, tempClusterChr = "" (:118-120)
      else if (pc != mc) f = 1;
      else f = ((int64_t)(int32_t)(uint32_t)p - (int64_t)start[k]) < 5 ? 1u : 0u;   // :175
      if (pc > mc) counters[1] = 1;                    // contig order went backwards: not coordinate sorted
    }
    flag[k] = f;
  }
}

struct ClusterAcc {      // device-side cluster record under construction
  unsigned long long first_read;   // min ordinal
  unsigned long long mask51;
  uint32_t num_reads;
  uint32_t num_t2c;
  uint32_t minus_members;          // minus-strand reads (all of them; the first is subtracted on the host)
  int32_t end;
};

struct PlParams2 {
  DeviceBatch b;
  PlRead rd;
  const uint32_t* cidx_incl;   // inclusive sum of flags: cluster number (1-based) of each kept read; 0 = head partial
  const uint32_t* ev_off;      // exclusive sum of nev
  ClusterAcc* cl;              // [n_clusters + 1], slot 0 = head partial
  unsigned long long* ev_key;  // (cluster << 32) | pos
  unsigned long long* ev_val;  // order key (ordinal << 6) | i
  uint64_t n;
};

// per-read contributions to the cluster records; runs of equal cluster inside a warp are combined first
__global__ void __launch_bounds__(256) pl_cluster_kernel(const PlParams2 P) {
  const uint32_t lane = threadIdx.x & 31;
  for (uint64_t base = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) & ~31ull; base < P.n;
       base += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = base + lane;
    const bool in = k < P.n;
    const unsigned long long key = in ? P.rd.key[k] : 0ull;
    const bool kept = key != 0;
    // non-kept records inherit the running cluster number with an empty contribution, so c is non-decreasing
    // across the warp and equality at distance d means the whole span is one cluster
    const uint32_t c = in ? P.cidx_incl[k] : 0xFFFFFFFFu;
    unsigned long long mask = kept ? P.rd.t2c[k] : 0ull;
    uint32_t reads = kept ? 1u : 0u;
    uint32_t t2c = __popcll(mask);
    uint32_t minus = (kept && (PS_META_FLAGS(P.b.meta[k]) & PS_RF_REVERSE)) ? 1u : 0u;
    int32_t end = kept ? P.rd.end[k] : INT32_MIN;
    unsigned long long first = kept ? k : ~0ull;
    // segmented (by cluster) inclusive scan over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t oc = __shfl_up_sync(0xFFFFFFFFu, c, d);
      const unsigned long long om = __shfl_up_sync(0xFFFFFFFFu, mask, d);
      const uint32_t orr = __shfl_up_sync(0xFFFFFFFFu, reads, d);
      const uint32_t ot = __shfl_up_sync(0xFFFFFFFFu, t2c, d);
      const uint32_t omi = __shfl_up_sync(0xFFFFFFFFu, minus, d);
      const int32_t oe = __shfl_up_sync(0xFFFFFFFFu, end, d);
      const unsigned long long of = __shfl_up_sync(0xFFFFFFFFu, first, d);
      if (lane >= (uint32_t)d && oc == c) {
        mask |= om; reads += orr; t2c += ot; minus += omi; end = oe > end ? oe : end; first = of < first ? of : first;
      }
    }
    // the last lane of each run publishes
    const uint32_t nc = __shfl_down_sync(0xFFFFFFFFu, c, 1);
    const bool tail = in && (lane == 31 || nc != c);
    if (tail && reads) {
      ClusterAcc* a = P.cl + c;
      atomicAdd(&a->num_reads, reads);
      if (t2c) atomicAdd(&a->num_t2c, t2c);
      if (minus) atomicAdd(&a->minus_members, minus);
      if (mask) atomicOr(&a->mask51, mask);
      atomicMax(&a->end, end);
      atomicMin(&a->first_read, first);
    }
    // T>C events
    if (kept) {
      unsigned long long m = P.rd.t2c[k];
      if (m) {
        const bool rev = PS_META_FLAGS(P.b.meta[k]) & PS_RF_REVERSE;
        const int32_t s = P.rd.start[k], e = P.rd.end[k];
        uint32_t o = P.ev_off[k];
        while (m) {
          const int i = __ffsll((long long)m) - 1;
          m &= m - 1;
          const int32_t pos = rev ? e - i : s + i;
          P.ev_key[o] = ((unsigned long long)c << 32) | (uint32_t)pos;
          P.ev_val[o] = ((unsigned long long)k << 6) | (unsigned)i;
          ++o;
        }
      }
    }
  }
}

// heads of runs of equal (cluster,pos) in the sorted event list
__global__ void pl_head_kernel(uint64_t n_ev, const unsigned long long* key, uint32_t* head) {
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n_ev; k += (uint64_t)gridDim.x * blockDim.x)
    head[k] = (k == 0 || key[k] != key[k - 1]) ? 1u : 0u;
}

struct SiteOut {
  unsigned long long key;        // (cluster << 32) | pos
  unsigned long long order_key;
  uint32_t t2c;
  uint32_t cov;
};

// one thread per site: run length, min order key, coverage = reads of the cluster whose [lo,hi] holds pos
__global__ void pl_site_kernel(uint64_t n_ev, const unsigned long long* key, const unsigned long long* val,
                               const uint32_t* head, const uint32_t* site_idx_excl, SiteOut* out,
                               const ClusterAcc* cl, const uint32_t* cluster_nreads_span /* unused */,
                               const unsigned long long* rkey, const uint32_t* cidx_incl, const int32_t* lo,
                               const int32_t* hi, const int32_t* start, uint64_t n_reads) {
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n_ev; k += (uint64_t)gridDim.x * blockDim.x) {
    if (!head[k]) continue;
    const unsigned long long ky = key[k];
    unsigned long long ok = val[k];
    uint32_t cnt = 1;
    for (uint64_t j = k + 1; j < n_ev && key[j] == ky; ++j) { ++cnt; ok = val[j] < ok ? val[j] : ok; }
    const uint32_t c = (uint32_t)(ky >> 32);
    const int32_t pos = (int32_t)(uint32_t)ky;
    // reads of cluster c: ordinals from first_read on, while cidx == c (non-kept reads interleave)
    uint32_t cov = 0;
    for (uint64_t r = cl[c].first_read == ~0ull ? 0 : cl[c].first_read; r < n_reads; ++r) {
      if (!rkey[r]) continue;
      if (cidx_incl[r] != c) break;
      if (start[r] > pos) break;               // sorted by start; lo >= start, so nothing later covers pos
      cov += (lo[r] <= pos && pos <= hi[r]) ? 1u : 0u;
    }
    SiteOut s;
    s.key = ky; s.order_key = ok; s.t2c = cnt; s.cov = cov;
    out[site_idx_excl[k]] = s;
  }
}

__global__ void pl_init_clusters(ClusterAcc* cl, uint64_t n) {
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
    ClusterAcc a;
    a.first_read = ~0ull; a.mask51 = 0; a.num_reads = 0; a.num_t2c = 0; a.minus_members = 0; a.end = INT32_MIN;
    cl[k] = a;
  }
}

template <typename T>
T* buf(ps_ctx* ctx, int slot, size_t count, cudaError_t& err) {
  if (err != cudaSuccess) return nullptr;
  err = ctx->pl_scratch[slot].reserve(count * sizeof(T) + 64);
  return static_cast<T*>(ctx->pl_scratch[slot].p);
}

}  // namespace

static int run_pileup(ps_ctx* ctx, const DeviceBatch& b, const ps_pileup_opts* opts, cudaStream_t st, ps_pileup** out) {
  ps_pileup* H = new ps_pileup();
  *out = H;
  const uint64_t n = b.n_reads;
  H->counters.num_reads_processed = n;
  if (n == 0) return PS_OK;
  if (n >= 0xFFFFFFFFull) return set_error(ctx, PS_ERR_UNSUPPORTED, "pileup batch of >= 2^32 reads");
  cudaError_t err = cudaSuccess;
  PlRead rd;
  rd.key = buf<unsigned long long>(ctx, 0, n, err);
  rd.t2c = buf<unsigned long long>(ctx, 1, n, err);
  unsigned long long* E = buf<unsigned long long>(ctx, 2, n, err);
  int32_t* i32 = buf<int32_t>(ctx, 3, 4 * n, err);
  uint32_t* u32 = buf<uint32_t>(ctx, 4, 4 * n + 8, err);
  unsigned long long* small = buf<unsigned long long>(ctx, 5, 8, err);   // fault, counters[2]
  if (err != cudaSuccess) return cuda_fail(ctx, err, "pileup scratch");
  rd.start = i32; rd.end = i32 + n; rd.lo = i32 + 2 * n; rd.hi = i32 + 3 * n;
  rd.flag = u32; rd.nev = u32 + n;
  uint32_t* cidx = u32 + 2 * n;
  uint32_t* ev_off = u32 + 3 * n;
  PS_CUDA(ctx, cudaMemsetAsync(small, 0xFF, 8, st));
  PS_CUDA(ctx, cudaMemsetAsync(small + 1, 0, 16, st));

  PlParams P;
  P.b = b; P.ref = ctx->ref; P.rd = rd; P.fault = small; P.counters = small + 1;
  P.n_tiles = (uint32_t)((n + PS_TILE_READS - 1) / PS_TILE_READS);
  const uint32_t grid = std::min<uint32_t>(P.n_tiles, (uint32_t)ctx->sm_count * 8);
  timer_begin(ctx, st);
  pl_read_kernel<<<grid, PS_BLOCK_THREADS, 0, st>>>(P);
  ctx->launches++;
  PS_CUDA(ctx, cudaGetLastError());

  // CUB temp storage
  size_t t1 = 0, t2 = 0, t3 = 0;
  cub::DeviceScan::InclusiveScan(nullptr, t1, rd.key, E, MaxOp(), (int64_t)n, st);
  cub::DeviceScan::InclusiveSum(nullptr, t2, rd.flag, cidx, (int64_t)n, st);
  cub::DeviceScan::ExclusiveSum(nullptr, t3, rd.nev, ev_off, (int64_t)n, st);
  size_t tmp_bytes = std::max(t1, std::max(t2, t3));
  void* tmp = buf<unsigned char>(ctx, 6, tmp_bytes, err);
  if (err != cudaSuccess) return cuda_fail(ctx, err, "cub temp");
  PS_CUDA(ctx, cub::DeviceScan::InclusiveScan(tmp, tmp_bytes, rd.key, E, MaxOp(), (int64_t)n, st));
  unsigned long long carry_key = 0;
  if (opts && opts->carry_valid)
    carry_key = ((unsigned long long)(opts->carry_contig + 1) << 32) | (uint32_t)opts->carry_cluster_end;
  const uint32_t g2 = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
  pl_flag_kernel<<<g2, 256, 0, st>>>(n, rd.key, E, rd.start, rd.flag, carry_key, small + 1);
  PS_CUDA(ctx, cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, rd.flag, cidx, (int64_t)n, st));
  PS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, rd.nev, ev_off, (int64_t)n, st));
  ctx->launches += 4;

  // totals
  uint32_t h_ncl = 0, h_lastoff = 0, h_lastnev = 0;
  unsigned long long h_small[3];
  PS_CUDA(ctx, cudaMemcpyAsync(&h_ncl, cidx + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaMemcpyAsync(&h_lastoff, ev_off + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaMemcpyAsync(&h_lastnev, rd.nev + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaMemcpyAsync(h_small, small, 24, cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  H->counters.skipped_due_indel = h_small[1];
  if (h_small[0] != PS_FAULT_NONE) {
    H->fault.code = (int32_t)(h_small[0] & 0xFF);
    H->fault.read_ordinal = h_small[0] >> 8;
    timer_end(ctx, st);
    return set_error(ctx, PS_ERR_REFERENCE_WOULD_THROW, "pileup: the JVM would die on a record of this batch");
  }
  if (h_small[2]) { timer_end(ctx, st); return set_error(ctx, PS_ERR_UNSORTED, ps_strerror(PS_ERR_UNSORTED)); }
  const uint64_t n_slots = (uint64_t)h_ncl + 1;          // slot 0 = reads continuing the carry-in cluster
  const uint64_t n_ev = (uint64_t)h_lastoff + h_lastnev;

  ClusterAcc* cl = buf<ClusterAcc>(ctx, 7, n_slots, err);
  unsigned long long* ev_key = buf<unsigned long long>(ctx, 8, 2 * n_ev + 2, err);
  unsigned long long* ev_val = buf<unsigned long long>(ctx, 9, 2 * n_ev + 2, err);
  uint32_t* ev_u32 = buf<uint32_t>(ctx, 10, 2 * n_ev + 2, err);
  if (err != cudaSuccess) return cuda_fail(ctx, err, "pileup cluster scratch");
  pl_init_clusters<<<(uint32_t)std::min<uint64_t>((n_slots + 255) / 256, 4096), 256, 0, st>>>(cl, n_slots);
  PlParams2 Q;
  Q.b = b; Q.rd = rd; Q.cidx_incl = cidx; Q.ev_off = ev_off; Q.cl = cl; Q.ev_key = ev_key; Q.ev_val = ev_val; Q.n = n;
  pl_cluster_kernel<<<g2, 256, 0, st>>>(Q);
  ctx->launches += 2;
  PS_CUDA(ctx, cudaGetLastError());

  uint64_t n_sites = 0;
  std::vector<SiteOut> h_sites;
  if (n_ev) {
    unsigned long long* sk = ev_key + n_ev;   // sorted keys / values
    unsigned long long* sv = ev_val + n_ev;
    size_t t4 = 0, t5 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t4, ev_key, sk, ev_val, sv, (int64_t)n_ev, 0, 64, st);
    uint32_t* head = ev_u32;
    uint32_t* sidx = ev_u32 + n_ev;
    cub::DeviceScan::ExclusiveSum(nullptr, t5, head, sidx, (int64_t)n_ev, st);
    size_t tb = std::max(t4, t5);
    void* tmp2 = buf<unsigned char>(ctx, 6, std::max(tb, tmp_bytes), err);
    if (err != cudaSuccess) return cuda_fail(ctx, err, "cub temp 2");
    PS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp2, tb, ev_key, sk, ev_val, sv, (int64_t)n_ev, 0, 64, st));
    const uint32_t g3 = (uint32_t)std::min<uint64_t>((n_ev + 255) / 256, (uint64_t)ctx->sm_count * 16);
    pl_head_kernel<<<g3, 256, 0, st>>>(n_ev, sk, head);
    PS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp2, tb, head, sidx, (int64_t)n_ev, st));
    uint32_t lh = 0, ls = 0;
    PS_CUDA(ctx, cudaMemcpyAsync(&lh, head + (n_ev - 1), 4, cudaMemcpyDeviceToHost, st));
    PS_CUDA(ctx, cudaMemcpyAsync(&ls, sidx + (n_ev - 1), 4, cudaMemcpyDeviceToHost, st));
    PS_CUDA(ctx, cudaStreamSynchronize(st));
    n_sites = (uint64_t)lh + ls;
    SiteOut* d_sites = buf<SiteOut>(ctx, 11, n_sites, err);
    if (err != cudaSuccess) return cuda_fail(ctx, err, "site buffer");
    pl_site_kernel<<<g3, 256, 0, st>>>(n_ev, sk, sv, head, sidx, d_sites, cl, nullptr, rd.key, cidx, rd.lo, rd.hi,
                                       rd.start, n);
    ctx->launches += 4;
    PS_CUDA(ctx, cudaGetLastError());
    h_sites.resize(n_sites);
    PS_CUDA(ctx, cudaMemcpyAsync(h_sites.data(), d_sites, n_sites * sizeof(SiteOut), cudaMemcpyDeviceToHost, st));
  }
  timer_end(ctx, st);
  std::vector<ClusterAcc> h_cl(n_slots);
  PS_CUDA(ctx, cudaMemcpyAsync(h_cl.data(), cl, n_slots * sizeof(ClusterAcc), cudaMemcpyDeviceToHost, st));
  // first-read attributes (start, strand, contig) of every cluster: gather on the host from small D2H reads
  std::vector<int32_t> h_start_all;   // only starts of first reads are needed: gather by kernel would be nicer
  PS_CUDA(ctx, cudaStreamSynchronize(st));

  // ---- assemble host records ------------------------------------------------------------------------
  // first-read data: fetch start / key / meta for each cluster's first read (one strided D2H gather each)
  std::vector<uint64_t> firsts(n_slots);
  for (uint64_t c = 0; c < n_slots; ++c) firsts[c] = h_cl[c].first_read;
  std::vector<int32_t> f_start(n_slots, 0);
  std::vector<unsigned long long> f_key(n_slots, 0);
  std::vector<uint32_t> f_meta(n_slots, 0);
  {
    // pull the whole start/key/meta arrays only when clusters are dense; otherwise element-wise copies would
    // dominate.  Reads are few bytes each, so a bulk copy is the simple choice here.
    std::vector<int32_t> a_start(n);
    std::vector<unsigned long long> a_key(n);
    std::vector<uint32_t> a_meta(n);
    PS_CUDA(ctx, cudaMemcpy(a_start.data(), rd.start, n * 4, cudaMemcpyDeviceToHost));
    PS_CUDA(ctx, cudaMemcpy(a_key.data(), rd.key, n * 8, cudaMemcpyDeviceToHost));
    PS_CUDA(ctx, cudaMemcpy(a_meta.data(), b.meta, n * 4, cudaMemcpyDeviceToHost));
    for (uint64_t c = 0; c < n_slots; ++c)
      if (firsts[c] != ~0ull) { f_start[c] = a_start[firsts[c]]; f_key[c] = a_key[firsts[c]]; f_meta[c] = a_meta[firsts[c]]; }
  }
  const uint32_t first_id = opts ? opts->first_running_id : 1;
  auto make = [&](uint64_t c, ps_cluster& o) {
    const ClusterAcc& a = h_cl[c];
    std::memset(&o, 0, sizeof(o));
    o.first_read = a.first_read;
    o.running_id = first_id + (uint32_t)c;
    o.contig = (uint32_t)(f_key[c] >> 32) - 1;
    o.start = f_start[c];
    o.end = a.end;
    o.num_reads = a.num_reads;
    o.num_t2c = a.num_t2c;
    o.first_reverse = (PS_META_FLAGS(f_meta[c]) & PS_RF_REVERSE) ? 1 : 0;
    o.minus_after_first = a.minus_members - (c ? o.first_reverse : 0);
    // StrandOrientation state at flush (P6): minus-first -> "-"; plus-first -> "+/-" once a minus member came
    o.combined_strand = o.first_reverse ? 1 : (o.minus_after_first ? 2 : 0);
    o.mask51 = a.mask51;
  };
  // sites are sorted by (cluster,pos)
  std::vector<uint64_t> site_lo(n_slots + 1, 0);
  {
    uint64_t s = 0;
    for (uint64_t c = 0; c < n_slots; ++c) {
      site_lo[c] = s;
      while (s < n_sites && (uint32_t)(h_sites[s].key >> 32) == (uint32_t)c) ++s;
    }
    site_lo[n_slots] = s;
  }
  auto put_sites = [&](uint64_t c, std::vector<ps_site>& dst) {
    for (uint64_t s = site_lo[c]; s < site_lo[c + 1]; ++s) {
      ps_site o;
      o.pos = (int32_t)(uint32_t)h_sites[s].key;
      o.t2c = h_sites[s].t2c;
      o.cov = h_sites[s].cov;
      o.reserved = 0;
      o.order_key = h_sites[s].order_key;
      dst.push_back(o);
    }
  };
  uint64_t dstr = 0;
  // slot 0: reads that continue the carry-in cluster (only with carry_valid)
  if (h_cl[0].num_reads) {
    H->has_head = true;
    make(0, H->head_partial);
    H->head_partial.running_id = 0;
    H->head_partial.minus_after_first = h_cl[0].minus_members;
    H->head_partial.site_begin = 0;
    put_sites(0, H->head_sites);
    H->head_partial.site_end = H->head_sites.size();
  }
  const uint64_t n_real = n_slots - 1;
  for (uint64_t c = 1; c <= n_real; ++c) {
    ps_cluster o;
    make(c, o);
    if (!o.first_reverse) dstr += o.minus_after_first;            // doubleStranded++ (:494-498)
    if (c < n_real) {
      o.site_begin = H->sites.size();
      put_sites(c, H->sites);
      o.site_end = H->sites.size();
      H->clusters.push_back(o);
    } else {                                                     // the last cluster is never flushed (:528-529)
      o.site_begin = 0;
      put_sites(c, H->open_sites);
      o.site_end = H->open_sites.size();
      H->open_cluster = o;
      H->counters.has_open_cluster = 1;
    }
  }
  // dense coverage of the boundary clusters from the [lo,hi] intervals of their reads
  {
    auto dense = [&](uint64_t r0, uint64_t r1, uint32_t slot, std::vector<uint32_t>& cov, int32_t& pos0) -> int {
      if (r1 <= r0) return PS_OK;
      const uint64_t m = r1 - r0;
      std::vector<int32_t> lo(m), hi(m);
      std::vector<unsigned long long> ky(m);
      std::vector<uint32_t> ci(m);
      PS_CUDA(ctx, cudaMemcpy(lo.data(), rd.lo + r0, m * 4, cudaMemcpyDeviceToHost));
      PS_CUDA(ctx, cudaMemcpy(hi.data(), rd.hi + r0, m * 4, cudaMemcpyDeviceToHost));
      PS_CUDA(ctx, cudaMemcpy(ky.data(), rd.key + r0, m * 8, cudaMemcpyDeviceToHost));
      PS_CUDA(ctx, cudaMemcpy(ci.data(), cidx + r0, m * 4, cudaMemcpyDeviceToHost));
      int32_t mn = INT32_MAX, mx = INT32_MIN;
      for (uint64_t k = 0; k < m; ++k)
        if (ky[k] && ci[k] == slot && lo[k] <= hi[k]) { mn = std::min(mn, lo[k]); mx = std::max(mx, hi[k]); }
      if (mn > mx) return PS_OK;
      pos0 = mn;
      cov.assign((size_t)(mx - mn + 1), 0);
      for (uint64_t k = 0; k < m; ++k)
        if (ky[k] && ci[k] == slot)
          for (int32_t p = lo[k]; p <= hi[k]; ++p) cov[p - mn]++;
      return PS_OK;
    };
    if (H->has_head) {
      const uint64_t r1 = n_real ? h_cl[1].first_read : n;
      int rc = dense(0, r1, 0, H->head_cov, H->head_cov_pos0);
      if (rc) return rc;
    }
    if (H->counters.has_open_cluster) {
      int rc = dense(h_cl[n_real].first_read, n, (uint32_t)n_real, H->open_cov, H->open_cov_pos0);
      if (rc) return rc;
    }
  }
  H->counters.double_stranded = dstr;
  H->counters.n_clusters = H->clusters.size();
  H->counters.n_sites = H->sites.size();
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

extern "C" {

int ps_pileup_batch_device(ps_ctx* ctx, const ps_read_batch* b, const ps_pileup_opts* opts, void* stream,
                           ps_pileup** out) {
  if (!ctx || !b || !out) return PS_ERR_INVALID_ARG;
  *out = nullptr;
  if (!ctx->ref_loaded) return set_error(ctx, PS_ERR_STATE, "no reference loaded");
  cudaSetDevice(ctx->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  int rc = run_pileup(ctx, pl_view_of(b), opts, st, out);
  if (rc != PS_OK && rc != PS_ERR_REFERENCE_WOULD_THROW) { delete *out; *out = nullptr; }
  return rc;
}

int ps_pileup_batch(ps_ctx* ctx, const ps_read_batch* hb, const ps_pileup_opts* opts, ps_pileup** out) {
  if (!ctx || !hb || !out) return PS_ERR_INVALID_ARG;
  *out = nullptr;
  if (!ctx->ref_loaded) return set_error(ctx, PS_ERR_STATE, "no reference loaded");
  cudaSetDevice(ctx->device);
  if (hb->n_reads == 0) { *out = new ps_pileup(); return PS_OK; }
  StagedBatch* sb = nullptr;
  int rc = stage_batch(ctx, hb, &sb);
  if (rc) return rc;
  rc = run_pileup(ctx, sb->view, opts, ctx->stream, out);
  if (rc != PS_OK && rc != PS_ERR_REFERENCE_WOULD_THROW) { delete *out; *out = nullptr; }
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
  uint64_t n = 0, used = 0;
  for (uint64_t c = first; c < h->clusters.size() && n < max_clusters; ++c) {
    const ps_cluster& s = h->clusters[c];
    const uint64_t ns = s.site_end - s.site_begin;
    if (used + ns > max_sites) break;
    clusters[n] = s;
    clusters[n].site_begin = used;
    clusters[n].site_end = used + ns;
    if (ns) std::memcpy(sites + used, h->sites.data() + s.site_begin, ns * sizeof(ps_site));
    used += ns;
    ++n;
  }
  return (int64_t)n;
}

int ps_pileup_open_cluster(ps_pileup* h, ps_cluster* cluster, ps_site* sites, uint64_t max_sites) {
  if (!h || !cluster) return PS_ERR_INVALID_ARG;
  if (!h->counters.has_open_cluster) return 0;
  if (h->open_sites.size() > max_sites) return PS_ERR_INVALID_ARG;
  *cluster = h->open_cluster;
  if (!h->open_sites.empty()) std::memcpy(sites, h->open_sites.data(), h->open_sites.size() * sizeof(ps_site));
  return 1;
}

int ps_pileup_head_partial(ps_pileup* h, ps_cluster* cluster, ps_site* sites, uint64_t max_sites) {
  if (!h || !cluster) return PS_ERR_INVALID_ARG;
  if (!h->has_head) return 0;
  if (h->head_sites.size() > max_sites) return PS_ERR_INVALID_ARG;
  *cluster = h->head_partial;
  if (!h->head_sites.empty()) std::memcpy(sites, h->head_sites.data(), h->head_sites.size() * sizeof(ps_site));
  return 1;
}

int64_t ps_pileup_boundary_coverage(ps_pileup* h, int which, int32_t* first_pos, uint32_t* cov, uint64_t max) {
  if (!h || !first_pos) return PS_ERR_INVALID_ARG;
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

void ps_pileup_close(ps_pileup* h) { delete h; }

}  // extern "C"
