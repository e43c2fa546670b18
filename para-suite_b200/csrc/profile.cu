// K1: error-profile count kernel.  Replaces the loop ErrorProfiling.java:146-409 (reference:
// /root/reference/src/src/utils/errorprofile/ErrorProfiling.java) for one SoA batch.
//
// One thread per read, one tile (256 reads) per block iteration, persistent blocks.  Counts go to
// per-block shared-memory histograms and are flushed once per block with 64-bit global atomics into the
// accumulator vector (layout: internal.h ProfileLayout), which is what the multi-GPU all-reduce sums.
#include "device_common.cuh"

namespace {

struct ProfileParams {
  DeviceBatch b;
  DeviceRef ref;
  ProfileLayout lay;
  unsigned long long* acc;    // int64 accumulators (two's complement adds)
  unsigned long long* fault;
  uint64_t ordinal0;
  uint32_t n_tiles;
};

__device__ __forceinline__ void warp_count(bool pred, unsigned long long* ctr_smem) {
  unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(ctr_smem, (unsigned long long)__popc(m));
}

// Generic (every CIGAR) per-read walk; literal restatement of ErrorProfiling.java:168-408 on packed data.
// Shared memory: s_conv[max_len*16] u32 | s_q[32] u64 (qsum, qcnt) | s_ctr[8] u64 | scan scratch
__global__ void __launch_bounds__(PS_BLOCK_THREADS) profile_generic_kernel(const ProfileParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t max_len = P.lay.max_len;
  unsigned long long* s_q = reinterpret_cast<unsigned long long*>(smem_raw);  // [32]
  unsigned long long* s_ctr = s_q + 32;                                       // [8]
  uint64_t* s_scan = reinterpret_cast<uint64_t*>(s_ctr + 8);                  // [8]
  uint32_t* s_conv = reinterpret_cast<uint32_t*>(s_scan + 8);                 // [max_len*16]

  for (uint32_t k = threadIdx.x; k < max_len * 16; k += blockDim.x) s_conv[k] = 0;
  if (threadIdx.x < 48) s_q[threadIdx.x] = 0;   // s_q, s_ctr, s_scan are contiguous (48 words)
  __syncthreads();

  for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const uint64_t r = (uint64_t)tile * PS_TILE_READS + threadIdx.x;
    const bool in_range = r < P.b.n_reads;
    const uint32_t meta = in_range ? __ldg(P.b.meta + r) : 0;
    const ReadOffsets off = read_offsets(P.b, tile, r, meta, in_range, s_scan);
    const uint32_t flags = PS_META_FLAGS(meta), L = PS_META_LEN(meta), ncig = PS_META_NCIGAR(meta);
    const uint64_t ordinal = P.ordinal0 + r;

    // filters :155-166 (first match wins)
    const bool f_unm = in_range && (flags & PS_RF_UNMAPPED);
    const bool f_dup = in_range && !f_unm && (flags & PS_RF_DUPLICATE);
    const bool f_zero = in_range && !f_unm && !f_dup && (flags & PS_RF_POS_ZERO);
    warp_count(f_unm, &s_ctr[PS_PC_UNMAPPED]);
    warp_count(f_dup, &s_ctr[PS_PC_DUPLICATES]);
    warp_count(f_zero, &s_ctr[PS_PC_START_ZERO]);
    bool live = in_range && !f_unm && !f_dup && !f_zero;

    const uint32_t* cig = P.b.cigar + off.cigar;
    uint32_t R = 0;
    bool has_indel = false;
    if (live) {
      for (uint32_t e = 0; e < ncig; ++e) {
        uint32_t c = __ldg(cig + e), op = c & 15u;
        if (op_consumes_ref(op)) R += c >> 4;
        has_indel |= (op == 1u) | (op == 2u);
      }
    }
    const uint64_t g0 = in_range ? __ldg(P.b.ref_start + r) : 0;
    // :169-172 FASTA fetch range
    if (live) {
      bool bad = (flags & PS_RF_REF_RANGE) || g0 >= P.ref.n_bases;
      if (!bad) {
        uint32_t c = contig_of(P.ref, g0);
        bad = g0 + R > __ldg(P.ref.contig_off + c + 1);
      }
      if (bad) { raise_fault(P.fault, ordinal, PS_THROW_REF_RANGE); live = false; }
    }
    warp_count(live, &s_ctr[PS_PC_NUM_READS_PROCESSED]);                       // :174
    if (live && R == 0) { raise_fault(P.fault, ordinal, PS_THROW_EMPTY_REF); live = false; }

    const uint32_t ml = L > R ? L : R;
    const bool walked = live && (L != R);
    bool skip = false;
    if (walked) {   // :194-299, pass 1: bounds (skip), indel side effects, uncaught exceptions
      int64_t pr = 0, pq = 0, pm = 0;
      for (uint32_t e = 0; e < ncig && live; ++e) {
        const uint32_t c = __ldg(cig + e), op = c & 15u;
        const int64_t n = c >> 4;
        if (op_is_match(op)) {
          // any z with z+pm >= ml, z+pr >= R or z+pq >= L throws inside the try -> skip
          if (n > 0 && (pm + n > (int64_t)ml || pr + n > (int64_t)R || pq + n > (int64_t)L)) skip = true;
          pm += n; pr += n; pq += n;
        } else if (op == 3u) {
          pr += n; pq += n;
        } else if (op == 1u || op == 2u) {
          if (n > 0 && pm + n > (int64_t)ml) { raise_fault(P.fault, ordinal, PS_THROW_INDEL_FILL); live = false; break; }
          pm += n;
          if (op == 1u) pq += n; else pr += n;
          unsigned long long* arr = P.acc + (op == 1u ? P.lay.ins : P.lay.del);
          for (int64_t q = 1; q <= n; ++q) {
            if (pm + q >= (int64_t)max_len) { raise_fault(P.fault, ordinal, PS_THROW_INDEL_POS); live = false; break; }
            atomicAdd(arr + (pm + q), 1ull);
          }
          if (!live) break;
          if (n > 1) atomicAdd(&s_ctr[PS_PC_LONGER_INDELS], 1ull);
        }
      }
      if (live) atomicAdd(&s_ctr[PS_PC_INDEL_READ], 1ull);                     // :296
    }
    if (live && skip) { atomicAdd(&s_ctr[PS_PC_SKIPPED_READS], 1ull); live = false; }  // :303-306

    if (live) {   // count loop :349-408 over the (virtual) temp arrays
      const bool rev = flags & PS_RF_REVERSE;
      const bool has_inv = flags & PS_RF_HAS_INVALID;
      const uint32_t qual_len = (flags & PS_RF_QUAL_MISSING) ? 0u : L;
      const uint8_t* rb = P.b.bases2 + off.base;
      const uint8_t* rq = P.b.qual + off.qual;
      const uint32_t rit = threadIdx.x;
      long long q_acc[4] = {0, 0, 0, 0};
      uint32_t q_cnt[4] = {0, 0, 0, 0};
      uint32_t checked = 0;
      uint32_t f_i = 0xFFFFFFFFu, f_code = 0;   // first uncaught exception of the count loop (by position i)
      // walk the M segments in the order that makes the final position i ascend
      int64_t pr = 0, pq = 0, pm = 0;
      // forward strand: columns ascend with the cigar; reverse strand: i = ml-1-col, so iterate ops backwards.
      // cursors at each op start are needed either way: recompute by a forward pass per op (ncig is small).
      const uint32_t n_ops = walked ? ncig : 1u;
      for (uint32_t step = 0; step < n_ops && live; ++step) {
        int64_t seg_col, seg_ref, seg_read, seg_len;
        if (!walked) {
          seg_col = 0; seg_ref = 0; seg_read = 0; seg_len = L;   // L == R: ungapped compare (Q1)
        } else {
          const uint32_t e_target = rev ? ncig - 1 - step : step;
          pr = pq = pm = 0;
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
          atomicAdd(&s_conv[i * 16 + a * 4 + b], 1u);
          ++checked;
          if (!has_indel) {
            if (i >= qual_len) { f_i = i; f_code = PS_THROW_QUAL_RANGE; live = false; break; }
            const long long qv = (long long)(signed char)__ldg(rq + i);   // qualities are NOT reversed (Q10)
            if (a == b) { q_acc[a] += qv; q_cnt[a]++; }
            else {
              atomicAdd(&s_q[a * 4 + b], (unsigned long long)qv);
              atomicAdd(&s_q[16 + a * 4 + b], 1ull);
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
      if (live) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (q_cnt[a]) {
            atomicAdd(&s_q[a * 5], (unsigned long long)q_acc[a]);
            atomicAdd(&s_q[16 + a * 5], (unsigned long long)q_cnt[a]);
          }
        atomicAdd(&s_ctr[PS_PC_TOTAL_BASES_CHECKED], (unsigned long long)checked);
        if (P.lay.infer_q)
          for (uint32_t i = 0; i < ml; ++i)
            atomicAdd(P.acc + P.lay.qhist + (size_t)i * 256 + __ldg(rq + i), 1ull);
      }
    }
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < max_len * 16; k += blockDim.x)
    if (s_conv[k]) atomicAdd(P.acc + P.lay.conv + k, (unsigned long long)s_conv[k]);
  if (threadIdx.x < 32 && s_q[threadIdx.x]) atomicAdd(P.acc + P.lay.qsum + threadIdx.x, s_q[threadIdx.x]);
  if (threadIdx.x < PS_PC_COUNT && s_ctr[threadIdx.x]) atomicAdd(P.acc + P.lay.ctr + threadIdx.x, s_ctr[threadIdx.x]);
}

}  // namespace

cudaError_t launch_profile(ps_ctx* ctx, const DeviceBatch& b, uint64_t ordinal0, cudaStream_t stream) {
  if (b.n_reads == 0) return cudaSuccess;
  ProfileParams P;
  P.b = b;
  P.ref = ctx->ref;
  P.lay = ctx->layout;
  P.acc = static_cast<unsigned long long*>(ctx->acc.p);
  P.fault = static_cast<unsigned long long*>(ctx->fault.p);
  P.ordinal0 = ordinal0;
  P.n_tiles = (uint32_t)((b.n_reads + PS_TILE_READS - 1) / PS_TILE_READS);
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
