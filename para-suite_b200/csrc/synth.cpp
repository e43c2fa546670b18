// Seeded synthetic PAR-CLIP workload generator.
//
// Not part of the drop-in library: it stands in for "a coordinate-sorted BAM + FASTA" when the named
// workloads are benchmarked (there is no network for datasets).  It writes the same packed reference and
// SoA read batch (include/parasuite_b200.h) the BAM batcher produces, so the GPU path and the CPU oracle
// read identical bytes.  Everything is a pure function of (seed, index) through a counter-based RNG, so
// any shard can regenerate any read range.
//
// Read model (a restatement of the *shape* of bin/createSimulatedPARCLIPDataset.pl, not its code):
// clusters of max(1, floor(N(16,10))) reads (pl:270) laid out left to right over the reference, start
// jitter +-3, one strand per cluster, 60% of clusters "bound" with 1-4 T sites converting T>C at rates
// .66/.24/.08/.04 (examples/simulation/example.sitefrequency), ~1% substitution errors, qualities
// clip(floor(N(mu_j, 4.2)), 3, 64) with mu_j = 31 falling to 25 over the last cycles, 0.1% N calls.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/parasuite_b200.h"

namespace {

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// counter-based stream: value k of stream (seed, a, b)
struct Rng {
  uint64_t key, ctr = 0;
  Rng(uint64_t seed, uint64_t a, uint64_t b) { key = splitmix64(seed ^ splitmix64(a * 0xD6E8FEB86659FD93ull + b)); }
  uint64_t next() { return splitmix64(key + (ctr++) * 0x9E3779B97F4A7C15ull); }
  double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
  double normal() {  // Box-Muller (cluster-level draws only)
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
  double normal_fast() {  // Irwin-Hall(4) ~ N(0,1), per-base draws
    uint64_t x = next();
    double s = (double)(x & 0xFFFF) + (double)((x >> 16) & 0xFFFF) + (double)((x >> 32) & 0xFFFF) + (double)(x >> 48);
    return (s * (1.0 / 65536.0) - 2.0) * 1.7320508075688772;
  }
};

// substitution model P(observed | true), rows/cols A,C,G,T (cumulative thresholds in 1/65536)
const double kErr[4][4] = {{0.990, 0.004, 0.003, 0.003},
                           {0.004, 0.990, 0.003, 0.003},
                           {0.006, 0.010, 0.977, 0.007},
                           {0.005, 0.005, 0.003, 0.987}};
const double kSiteRate[4] = {0.66, 0.24, 0.08, 0.04};

struct Synth {
  ps_read_batch batch;
  std::vector<uint32_t> meta, ref_start, cigar, tile_exc_off, exc;
  std::vector<uint8_t> bases2, qual;
  std::vector<uint64_t> tile_base_off, tile_qual_off, tile_cigar_off;
  uint64_t n_clusters = 0;
};

inline uint32_t ref_code(const ps_reference* ref, uint64_t g) { return (ref->seq2[g >> 4] >> (2 * (g & 15))) & 3; }
inline bool ref_inv(const ps_reference* ref, uint64_t g) { return (ref->inv[g >> 5] >> (g & 31)) & 1; }

struct Cluster {
  uint64_t base;      // global 0-based position of the nominal start
  uint32_t n_reads;
  bool minus, bound;
};

}  // namespace

extern "C" {

typedef struct ps_synth_params {
  uint64_t seed;
  uint64_t n_reads;
  uint32_t read_len;
  uint32_t mode;         /* 0: one M op per read; 1: dense CIGAR (soft clips + indels), config 5 */
  uint32_t threads;
  uint32_t special_ppm;  /* per-million rate of duplicate / unmapped / POS==0 records (0 for the benches) */
  uint64_t region_lo;    /* clusters are laid out over global offsets [region_lo, region_hi) */
  uint64_t region_hi;
  uint32_t n_ppm;        /* per-million rate of N base calls in reads (default 1000) */
  uint32_t reserved;
} ps_synth_params;

// Reference: contigs of the given lengths, iid uniform ACGT, ~1% of bases in N runs of `n_run` bases.
// seq2 / inv must hold ceil(n/16)+8 and ceil(n/32)+8 words.
int ps_synth_reference(uint64_t seed, uint32_t n_contigs, const uint64_t* contig_len, uint32_t n_run, uint32_t* seq2,
                       uint32_t* inv, int threads) {
  uint64_t n = 0;
  for (uint32_t c = 0; c < n_contigs; ++c) n += contig_len[c];
  uint64_t w2 = (n + 15) / 16 + 8, w1 = (n + 31) / 32 + 8;
  std::memset(inv, 0, w1 * 4);
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([=]() {
      for (uint64_t w = w2 * t / threads; w < w2 * (t + 1) / threads; ++w) {
        uint64_t x = splitmix64(seed * 0x2545F4914F6CDD1Dull + w);
        seq2[w] = (uint32_t)(x >> 16);
      }
    });
  for (auto& th : pool) th.join();
  // zero the tail beyond n
  for (uint64_t g = n; g < w2 * 16; ++g) seq2[g >> 4] &= ~(3u << (2 * (g & 15)));
  if (n_run) {
    uint64_t runs = n / (100ull * n_run);
    Rng rng(seed, 0xABCD, 1);
    for (uint64_t k = 0; k < runs; ++k) {
      uint64_t a = (uint64_t)(rng.uni() * (double)(n > n_run ? n - n_run : 0));
      for (uint64_t g = a; g < a + n_run && g < n; ++g) {
        inv[g >> 5] |= 1u << (g & 31);
        seq2[g >> 4] &= ~(3u << (2 * (g & 15)));
      }
    }
  }
  return PS_OK;
}

void ps_synth_free(Synth* s) { delete s; }
const ps_read_batch* ps_synth_batch(const Synth* s) { return &s->batch; }
uint64_t ps_synth_n_clusters(const Synth* s) { return s->n_clusters; }

Synth* ps_synth_reads(const ps_synth_params* P, const ps_reference* ref) {
  const uint64_t n = P->n_reads;
  const uint32_t L = P->read_len;
  const uint32_t T = PS_TILE_READS;
  const uint32_t max_span = L + 16;  // reference span incl. jitter and deletions
  uint64_t lo = P->region_lo, hi = P->region_hi ? P->region_hi : ref->n_bases;
  if (hi > ref->n_bases) hi = ref->n_bases;
  if (n == 0 || L == 0 || L > 4000 || hi <= lo + 4 * max_span) return nullptr;
  int threads = P->threads ? (int)P->threads : 1;
  uint32_t n_ppm = P->n_ppm;

  // ---- clusters: sizes until the read count is reached ---------------------------------------------
  std::vector<Cluster> cl;
  cl.reserve(n / 12 + 16);
  std::vector<uint64_t> first;  // first read of each cluster
  uint64_t total = 0;
  for (uint64_t c = 0; total < n; ++c) {
    Rng r(P->seed, 1, c);
    double z = 16.0 + 10.0 * r.normal();
    uint32_t k = z < 1.0 ? 1u : (uint32_t)z;
    if (total + k > n) k = (uint32_t)(n - total);
    Cluster x;
    x.n_reads = k;
    x.minus = r.next() & 1;
    x.bound = r.uni() < 0.6;
    x.base = 0;
    cl.push_back(x);
    first.push_back(total);
    total += k;
  }
  const uint64_t nc = cl.size();
  // ---- positions: clusters left to right, never straddling a contig boundary -----------------------
  {
    // usable length per contig inside [lo, hi)
    std::vector<uint64_t> cbeg, cend;
    uint64_t usable = 0;
    for (uint32_t k = 0; k < ref->n_contigs; ++k) {
      uint64_t a = std::max(lo, ref->contig_off[k]) + 4, b = std::min(hi, ref->contig_off[k + 1]);
      if (b > a + 2 * max_span) { cbeg.push_back(a); cend.push_back(b - max_span - 4); usable += cend.back() - a; }
    }
    if (!usable) return nullptr;
    double spacing = (double)usable / (double)nc;
    // walk: cluster c sits at usable-coordinate c*spacing + jitter
    size_t k = 0;
    uint64_t acc = 0;  // usable coordinate of cbeg[k]
    for (uint64_t c = 0; c < nc; ++c) {
      Rng r(P->seed, 2, c);
      double room = spacing > 56.0 ? spacing - 48.0 : (spacing > 8.0 ? spacing - 8.0 : 0.0);
      uint64_t u = (uint64_t)(c * spacing + r.uni() * room);
      while (k + 1 < cbeg.size() && u >= acc + (cend[k] - cbeg[k])) { acc += cend[k] - cbeg[k]; ++k; }
      uint64_t g = cbeg[k] + (u - acc);
      if (g >= cend[k]) g = cend[k] - 1;
      cl[c].base = g;
    }
  }

  Synth* S = new Synth();
  S->n_clusters = nc;
  const uint64_t n_tiles = (n + T - 1) / T;
  S->meta.assign(n, 0);
  S->ref_start.assign(n, 0);
  S->tile_base_off.assign(n_tiles + 1, 0);
  S->tile_qual_off.assign(n_tiles + 1, 0);
  S->tile_cigar_off.assign(n_tiles + 1, 0);
  S->tile_exc_off.assign(n_tiles + 1, 0);

  // ---- phase 1: per-read start, cigar, flags (keyed by cluster / slot) ------------------------------
  std::vector<uint8_t> ncig(n, 1);
  std::vector<uint32_t> cigtmp(P->mode ? n * 7 : 0);
  auto phase1 = [&](uint64_t c0, uint64_t c1) {
    std::vector<std::pair<uint64_t, uint32_t>> order;
    for (uint64_t c = c0; c < c1; ++c) {
      const Cluster& x = cl[c];
      order.resize(x.n_reads);
      for (uint32_t j = 0; j < x.n_reads; ++j) {
        Rng r(P->seed, 3, first[c] + j);
        order[j] = {x.base + 3 + r.below(7) - 3, j};
      }
      std::sort(order.begin(), order.end());
      for (uint32_t j = 0; j < x.n_reads; ++j) {
        uint64_t rd = first[c] + j;
        uint32_t fl = x.minus ? PS_RF_REVERSE : 0;
        Rng r(P->seed, 4, rd);
        if (P->special_ppm && r.below(1000000) < P->special_ppm) {
          uint32_t w = r.below(3);
          fl |= w == 0 ? PS_RF_DUPLICATE : (w == 1 ? PS_RF_UNMAPPED : PS_RF_POS_ZERO);
        }
        uint32_t k = 1;
        if (P->mode == 1) {
          // [aS] M (I|D) M (I|D) M [bS]; indels >= 8 nt from either end of the aligned part
          uint32_t a = r.uni() < 0.2 ? 1 + r.below(10) : 0, b = r.uni() < 0.2 ? 1 + r.below(10) : 0;
          uint32_t nind = r.below(3);
          uint32_t body = L - a - b;
          uint32_t* cg = &cigtmp[rd * 7];
          k = 0;
          if (a) cg[k++] = (a << 4) | 4;
          uint32_t left = body;
          for (uint32_t e = 0; e < nind && left > 40; ++e) {
            uint32_t m = 8 + r.below(left - 24 > 0 ? (left - 24) / (nind - e) : 1);
            uint32_t len = 1 + r.below(3);
            bool ins = r.next() & 1;
            cg[k++] = (m << 4) | 0;
            left -= m;
            if (ins) { cg[k++] = (len << 4) | 1; left -= len; }
            else cg[k++] = (len << 4) | 2;
          }
          cg[k++] = (left << 4) | 0;
          if (b) cg[k++] = (b << 4) | 4;
        }
        ncig[rd] = (uint8_t)k;
        S->ref_start[rd] = (uint32_t)order[j].first;
        S->meta[rd] = PS_MAKE_META(L, k, fl);
      }
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(phase1, nc * t / threads, nc * (t + 1) / threads);
    for (auto& th : pool) th.join();
  }
  // ---- coordinate order ------------------------------------------------------------------------------
  // A BAM is sorted by start throughout, not only inside a cluster: neighbouring clusters that overlap interleave.
  // slot[rd] = place of generator read rd in the batch (stable: ties keep generator order); contents stay keyed by rd.
  std::vector<uint64_t> slot(n);
  {
    std::vector<uint64_t> idx(n);
    for (uint64_t r = 0; r < n; ++r) idx[r] = r;
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return S->ref_start[a] < S->ref_start[b]; });
    std::vector<uint32_t> m2(n), s2(n);
    std::vector<uint8_t> c2(n);
    std::vector<uint32_t> g2(cigtmp.size());
    for (uint64_t k = 0; k < n; ++k) {
      const uint64_t r = idx[k];
      slot[r] = k;
      m2[k] = S->meta[r]; s2[k] = S->ref_start[r]; c2[k] = ncig[r];
      if (P->mode) std::memcpy(&g2[k * 7], &cigtmp[r * 7], 7 * 4);
    }
    S->meta.swap(m2); S->ref_start.swap(s2); ncig.swap(c2); cigtmp.swap(g2);
  }
  // ---- offsets ---------------------------------------------------------------------------------------
  const uint32_t bpr = (L + 3) / 4;
  std::vector<uint64_t> coff(n + 1, 0);
  for (uint64_t r = 0; r < n; ++r) coff[r + 1] = coff[r] + ncig[r];
  for (uint64_t t = 0; t <= n_tiles; ++t) {
    uint64_t r = std::min<uint64_t>(t * T, n);
    S->tile_base_off[t] = r * bpr;
    S->tile_qual_off[t] = r * (uint64_t)L;
    S->tile_cigar_off[t] = coff[r];
  }
  S->bases2.assign(n * bpr + 64, 0);
  S->qual.assign(n * (uint64_t)L + 64, 0);
  S->cigar.assign(coff[n] + 16, 0);

  // ---- phase 2: bases, qualities, cigar stream, N calls ----------------------------------------------
  std::vector<std::vector<uint32_t>> exc_by_thread(threads);   // (tile<<?) handled below: store global read + pos
  std::vector<std::vector<uint64_t>> exc_read_by_thread(threads);
  auto phase2 = [&](int tid, uint64_t c0, uint64_t c1) {
    std::vector<uint8_t> fwd(L);  // forward-strand codes of the read
    for (uint64_t c = c0; c < c1; ++c) {
      const Cluster& x = cl[c];
      // conversion sites of a bound cluster: genome positions holding T (plus) / A (minus)
      uint64_t site[4];
      uint32_t nsite = 0;
      if (x.bound) {
        Rng r(P->seed, 5, c);
        uint32_t want = 1 + r.below(4);
        uint32_t want_code = x.minus ? 0u : 3u;
        for (uint32_t s = 0; s < want; ++s) {
          uint64_t g = x.base + 3 + r.below(L);
          for (uint32_t tries = 0; tries < L; ++tries, ++g) {
            if (g >= x.base + 3 + L) g = x.base + 3;
            if (!ref_inv(ref, g) && ref_code(ref, g) == want_code) break;
          }
          site[nsite++] = g;
        }
      }
      for (uint32_t j = 0; j < x.n_reads; ++j) {
        const uint64_t gen = first[c] + j;       // generator ordinal: keys the random streams
        const uint64_t rd = slot[gen];           // place in the batch
        uint32_t m = S->meta[rd];
        bool minus = PS_META_FLAGS(m) & PS_RF_REVERSE;
        uint32_t k = ncig[rd];
        uint32_t one = (L << 4) | 0;
        const uint32_t* cg = P->mode ? &cigtmp[rd * 7] : &one;
        std::memcpy(&S->cigar[coff[rd]], cg, 4 * k);
        Rng r(P->seed, 6, gen);
        // true forward bases through the cigar
        uint64_t g = S->ref_start[rd];
        uint32_t p = 0;
        for (uint32_t e = 0; e < k; ++e) {
          uint32_t op = cg[e] & 15, len = cg[e] >> 4;
          if (op == 0) {
            for (uint32_t z = 0; z < len; ++z, ++g, ++p) {
              uint32_t code = ref_inv(ref, g) ? r.below(4) : ref_code(ref, g);
              // T>C conversion (read orientation): plus T->C, minus (forward view) A->G
              for (uint32_t s = 0; s < nsite; ++s)
                if (site[s] == g && r.uni() < kSiteRate[s] && !ref_inv(ref, g)) code = minus ? 2u : 1u;
              fwd[p] = (uint8_t)code;
            }
          } else if (op == 1 || op == 4) {
            for (uint32_t z = 0; z < len; ++z, ++p) fwd[p] = (uint8_t)r.below(4);
          } else if (op == 2) {
            g += len;
          }
        }
        // sequencing errors + N calls in read orientation; qualities by cycle
        uint8_t* q = &S->qual[rd * (uint64_t)L];
        uint8_t* b2 = &S->bases2[rd * (uint64_t)bpr];
        bool any_n = false;
        for (uint32_t cyc = 0; cyc < L; ++cyc) {
          uint32_t fp = minus ? L - 1 - cyc : cyc;  // forward index of sequencing cycle `cyc`
          uint32_t t = minus ? 3u - fwd[fp] : fwd[fp];
          double u = r.uni(), acc = 0;
          uint32_t obs = 3;
          for (uint32_t o = 0; o < 4; ++o) { acc += kErr[t][o]; if (u < acc) { obs = o; break; } }
          uint32_t fobs = minus ? 3u - obs : obs;
          double mu = cyc + 5 >= L ? 31.0 - (double)(cyc + 6 - L) : 31.0;
          double qv = std::floor(mu + 4.2 * r.normal_fast());
          q[fp] = (uint8_t)(qv < 3 ? 3 : (qv > 64 ? 64 : qv));
          if (n_ppm && r.below(1000000) < n_ppm) {
            any_n = true;
            fobs = 0;
            exc_read_by_thread[tid].push_back(rd);
            exc_by_thread[tid].push_back(fp);
          }
          b2[fp >> 2] |= (uint8_t)(fobs << (2 * (fp & 3)));
        }
        if (any_n) S->meta[rd] |= PS_RF_HAS_INVALID << 24;
      }
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(phase2, t, nc * t / threads, nc * (t + 1) / threads);
    for (auto& th : pool) th.join();
  }
  // exceptions: ordered by (read, position)
  {
    std::vector<std::pair<uint64_t, uint32_t>> all;
    for (int t = 0; t < threads; ++t)
      for (size_t k = 0; k < exc_by_thread[t].size(); ++k) all.push_back({exc_read_by_thread[t][k], exc_by_thread[t][k]});
    std::sort(all.begin(), all.end());
    S->exc.reserve(all.size() + 16);
    size_t k = 0;
    for (uint64_t t = 0; t < n_tiles; ++t) {
      S->tile_exc_off[t] = (uint32_t)S->exc.size();
      while (k < all.size() && all[k].first < (t + 1) * T) {
        S->exc.push_back((uint32_t)((all[k].first % T) << 16) | all[k].second);
        ++k;
      }
    }
    S->tile_exc_off[n_tiles] = (uint32_t)S->exc.size();
    S->exc.resize(S->exc.size() + 16, 0);
  }
  ps_read_batch& B = S->batch;
  std::memset(&B, 0, sizeof(B));
  B.n_reads = n;
  B.meta = S->meta.data();
  B.ref_start = S->ref_start.data();
  B.bases2 = S->bases2.data();
  B.qual = S->qual.data();
  B.cigar = S->cigar.data();
  B.tile_base_off = S->tile_base_off.data();
  B.tile_qual_off = S->tile_qual_off.data();
  B.tile_cigar_off = S->tile_cigar_off.data();
  B.tile_exc_off = S->tile_exc_off.data();
  B.exc = S->exc.data();
  B.uniform_len = L;
  B.uniform_ncigar = P->mode == 0 ? 1 : 0;
  B.bases_bytes = n * bpr;
  B.qual_bytes = n * (uint64_t)L;
  B.cigar_count = coff[n];
  B.exc_count = S->tile_exc_off[n_tiles];
  B.max_len = L;
  return S;
}

}  // extern "C"
