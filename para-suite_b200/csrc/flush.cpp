// Host side of the cluster flush: what PileupClusters.java does with a closed cluster before it writes its rows
// (reference: /root/reference/src/src/utils/pileupclusters/PileupClusters.java:178-260, end of run :529-545;
// SNPCalling.java:49-69), fed with the cluster / site records the pileup kernels return.  Host code, no GPU needed.
//
// Everything here is sequential by construction in the reference -- running sums over clusters in file order, and an
// anchor-site tie-break that depends on the iteration order of ONE java.util.HashMap that lives for the whole run
// (cleared, never re-created, so its capacity is the historical maximum).  That map is modelled below (JDK 8+:
// hash = h ^ h>>>16, power-of-two table starting at 16, tail insertion, doubling at 0.75 load with order-preserving
// split, clear() keeps the table, putAll() pre-sizes an unallocated table).  Buckets that would be treeified
// (9 keys in one bucket) are reported as unsupported instead of being guessed.
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "parasuite_b200.h"

namespace {

// std::stable_sort asks for a temporary buffer at every call; a cluster has a handful of sites
template <typename It, typename Cmp>
void stable_sort_small(It a, It b, Cmp less) {
  if (b - a > 32) { std::stable_sort(a, b, less); return; }
  for (It i = a == b ? b : a + 1; i < b; ++i) {
    auto v = *i;
    It j = i;
    while (j > a && less(v, *(j - 1))) { *j = *(j - 1); --j; }
    *j = v;
  }
}

struct JMap {   // java.util.HashMap<Integer,Integer>, iteration order included
  // Buckets are fixed arrays (a ninth key in one bucket is "unsupported" anyway) in one flat table, and the buckets in use
  // are listed, so that clearing and walking the map costs what the cluster's few sites cost, not the table's size, and
  // nothing is allocated per cluster.
  struct Node { int32_t k; int32_t v; };
  struct Bucket { uint32_t n; Node e[8]; };
  std::vector<Bucket> table;
  mutable std::vector<uint32_t> used;       // buckets that have held a node since the last clear (any order)
  size_t size = 0, threshold = 0;
  bool allocated = false, unsupported = false;

  static uint32_t hash(int32_t k) { const uint32_t h = (uint32_t)k; return h ^ (h >> 16); }
  static size_t table_size_for(size_t c) { size_t n = 1; while (n < c) n <<= 1; return std::min<size_t>(std::max<size_t>(n, 1), (size_t)1 << 30); }
  void wipe() {
    for (uint32_t j : used) table[j].n = 0;
    used.clear();
    size = 0;
  }
  void reset() {   // `new HashMap<>()`: no table yet (the storage is kept, all buckets empty)
    wipe();
    threshold = 0;
    allocated = false;
    unsupported = false;
  }
  void push(Bucket& b, uint32_t j, Node n) {
    if (b.n >= 8) { unsupported = true; return; }   // treeifyBin: link order no longer follows insertion
    if (b.n == 0) used.push_back(j);
    b.e[b.n++] = n;
  }
  void resize() {
    if (!allocated) {
      const size_t cap = threshold ? threshold : 16;
      if (table.size() != cap) table.assign(cap, Bucket{});
      threshold = cap * 3 / 4;
      allocated = true;
      return;
    }
    const size_t ocap = table.size(), ncap = ocap * 2;
    std::vector<Bucket> old(ncap);
    old.swap(table);
    std::vector<uint32_t> was;
    was.swap(used);
    for (uint32_t j : was) {
      const Bucket& b = old[j];
      for (uint32_t q = 0; q < b.n; ++q) {          // lo / hi split keeps order
        const uint32_t t = (hash(b.e[q].k) & ocap) ? j + (uint32_t)ocap : j;
        push(table[t], t, b.e[q]);
      }
    }
    threshold = ncap * 3 / 4;
  }
  void put(int32_t k, int32_t v) {
    if (!allocated) resize();
    const uint32_t j = hash(k) & (uint32_t)(table.size() - 1);
    Bucket& b = table[j];
    for (uint32_t q = 0; q < b.n; ++q)
      if (b.e[q].k == k) { b.e[q].v = v; return; }
    push(b, j, {k, v});
    if (++size > threshold) resize();
  }
  const int32_t* get(int32_t k) const {
    if (!allocated) return nullptr;
    const Bucket& b = table[hash(k) & (table.size() - 1)];
    for (uint32_t q = 0; q < b.n; ++q)
      if (b.e[q].k == k) return &b.e[q].v;
    return nullptr;
  }
  void remove(int32_t k) {
    if (!allocated) return;
    Bucket& b = table[hash(k) & (table.size() - 1)];
    for (uint32_t q = 0; q < b.n; ++q)
      if (b.e[q].k == k) {
        for (uint32_t z = q + 1; z < b.n; ++z) b.e[z - 1] = b.e[z];
        --b.n; --size;
        return;
      }
  }
  void clear() { wipe(); }   // HashMap.clear(): the table stays
  template <typename F>
  void for_each(F f) const {   // table order: bucket index, then link order
    std::sort(used.begin(), used.end());
    for (uint32_t j : used) {
      const Bucket& b = table[j];
      for (uint32_t q = 0; q < b.n; ++q) f(b.e[q].k, b.e[q].v);
    }
  }
  void put_all(const JMap& o) {   // HashMap.putMapEntries
    const size_t s = o.size;
    if (s == 0) return;
    if (!allocated) {
      const double ft = (double)s / 0.75 + 1.0;
      const size_t t = ft < (double)((size_t)1 << 30) ? (size_t)ft : (size_t)1 << 30;
      if (t > threshold) threshold = table_size_for(t);
    } else if (s > threshold) resize();
    o.for_each([&](int32_t k, int32_t v) { put(k, v); });
  }
};

}  // namespace

struct ps_flush {
  uint32_t min_cov = 1;
  std::vector<std::string> contig_query;   // contig name as SNPCalling.querySNP asks for it (leading "chr" stripped)
  std::unordered_map<std::string, std::unordered_set<int32_t>> snps;   // VCF rows with T in REF and C in ALT[0]
  JMap mutation_map;                        // the one map of the run (PileupClusters.java:126)
  JMap tmp;                                 // storage of the per-cluster copy (:187), reset for every cluster
  std::vector<std::pair<int32_t, uint32_t>> cov;
  ps_flush_totals totals{};
  std::vector<double> afi;                  // alleleFrequencyInformation
  std::string err;
};

extern "C" {

int ps_flush_create(ps_flush** out, uint32_t min_read_coverage, uint32_t n_contigs, const char* const* contig_names) {
  if (!out || (n_contigs && !contig_names)) return PS_ERR_INVALID_ARG;
  ps_flush* f = new ps_flush();
  f->min_cov = min_read_coverage;
  for (uint32_t c = 0; c < n_contigs; ++c) {
    std::string n = contig_names[c] ? contig_names[c] : "";
    if (n.rfind("chr", 0) == 0) n = n.substr(3);          // SNPCalling.java:51-54
    f->contig_query.push_back(n);
  }
  *out = f;
  return PS_OK;
}

void ps_flush_destroy(ps_flush* f) { delete f; }
const char* ps_flush_error(const ps_flush* f) { return f ? f->err.c_str() : "no object"; }

// one VCF row (CHROM, POS, REF, first ALT allele); kept only if it can ever answer querySNP(chr, pos, "T", "C")
int ps_flush_add_snp(ps_flush* f, const char* chrom, int32_t pos, const char* ref, const char* alt0) {
  if (!f || !chrom || !ref || !alt0) return PS_ERR_INVALID_ARG;
  if (strchr(ref, 'T') && strchr(alt0, 'C')) f->snps[chrom].insert(pos);   // String.contains, case sensitive (:61-64)
  return PS_OK;
}

int ps_flush_load_vcf(ps_flush* f, const char* path) {   // plain, gzip or bgzip (BGZF is multi-member gzip)
  if (!f || !path) return PS_ERR_INVALID_ARG;
  gzFile g = gzopen(path, "rb");
  if (!g) { f->err = std::string("cannot open ") + path; return PS_ERR_IO; }
  std::string line;
  char buf[1 << 16];
  while (gzgets(g, buf, sizeof buf)) {
    line += buf;
    if (line.empty() || line.back() != '\n') { if (!gzeof(g)) continue; }
    if (line[0] != '#') {
      std::vector<std::string> col;
      size_t a = 0;
      while (col.size() < 5) {
        const size_t b = line.find('\t', a);
        col.push_back(line.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
      }
      if (col.size() == 5) {
        while (!col[4].empty() && (col[4].back() == '\n' || col[4].back() == '\r')) col[4].pop_back();
        const std::string alt0 = col[4].substr(0, col[4].find(','));
        ps_flush_add_snp(f, col[0].c_str(), (int32_t)atol(col[1].c_str()), col[3].c_str(), alt0.c_str());
      }
    }
    line.clear();
  }
  gzclose(g);
  return PS_OK;
}

// The flush of `n` closed clusters, in order (PileupClusters.java:178-260).  sites[] is indexed by the clusters'
// site_begin / site_end (as ps_pileup_next returns them).  rows[k] describes cluster k.
int ps_flush_clusters(ps_flush* f, const ps_cluster* clusters, uint64_t n, const ps_site* sites, ps_flush_row* rows) {
  if (!f || (n && (!clusters || !rows))) return PS_ERR_INVALID_ARG;
  std::vector<const ps_site*> order;
  std::vector<double> sorted;
  for (uint64_t c = 0; c < n; ++c) {
    const ps_cluster& cl = clusters[c];
    ps_flush_row& row = rows[c];
    memset(&row, 0, sizeof row);
    row.best_pos = -1;
    if (cl.site_end > cl.site_begin && !sites) return PS_ERR_INVALID_ARG;
    // mutationMap as the record loop left it: cleared at the cluster's first read (:353), keys put in the order
    // their first T>C was seen (ps_site.order_key).  The loop fills the ONE map of the run for every cluster, also for
    // those below minReadCoverage: their puts can grow the table, and clear() keeps the grown capacity, which decides
    // the iteration order -- and with it the anchor tie-break -- of every later cluster.
    if (cl.site_end == cl.site_begin) {      // no T>C seen: the map is cleared and stays empty, every step below is a no-op
      f->mutation_map.clear();
      if (cl.num_reads >= f->min_cov) row.emitted = 1;
      continue;
    }
    order.clear();
    for (uint64_t s = cl.site_begin; s < cl.site_end; ++s) order.push_back(sites + s);
    if (order.size() > 1) std::sort(order.begin(), order.end(), [](const ps_site* a, const ps_site* b) { return a->order_key < b->order_key; });
    JMap& mm = f->mutation_map;
    mm.clear();
    std::vector<std::pair<int32_t, uint32_t>>& cov = f->cov;           // baseCoveredMap is only ever read at these keys
    cov.clear();
    for (const ps_site* s : order) { mm.put(s->pos, (int32_t)s->t2c); cov.push_back({s->pos, s->cov}); }
    stable_sort_small(cov.begin(), cov.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    auto cov_at = [&](int32_t k) {                                     // the last value put under k
      auto it = std::upper_bound(cov.begin(), cov.end(), k, [](int32_t key, const auto& e) { return key < e.first; });
      return it == cov.begin() ? 0u : (it - 1)->second;
    };
    if (mm.unsupported) {
      f->err = "a HashMap bucket would be treeified (9 T>C positions of one cluster in one bucket): iteration order not modelled";
      return PS_ERR_UNSUPPORTED;
    }
    if (cl.num_reads < f->min_cov) continue;                           // :180
    row.emitted = 1;
    row.num_t2c_sites = (uint32_t)mm.size;                             // :181 (before the SNP filter)
    // SNP filter :187-201
    JMap& tmp = f->tmp;
    tmp.reset();
    tmp.put_all(mm);
    const std::unordered_set<int32_t>* known = nullptr;
    if (cl.contig < f->contig_query.size()) {
      auto it = f->snps.find(f->contig_query[cl.contig]);
      if (it != f->snps.end()) known = &it->second;
    }
    mm.for_each([&](int32_t k, int32_t v) {
      if (known && known->count(k)) { tmp.remove(k); f->totals.snp_hit++; }
      if (v == 1) f->totals.high_frequent_error++;
    });
    mm.clear();
    mm.put_all(tmp);
    if (mm.unsupported || tmp.unsupported) {
      f->err = "a HashMap bucket would be treeified (9 T>C positions of one cluster in one bucket): iteration order not modelled";
      return PS_ERR_UNSUPPORTED;
    }
    double fraction = 0.0;
    if (row.num_t2c_sites > 0) {                                       // :203
      sorted.clear();
      double best = 0.0;
      int32_t best_pos = -1;
      mm.for_each([&](int32_t k, int32_t v) {
        const double val = (double)v / (double)cov_at(k);
        if (val >= best) { best = val; best_pos = k; }                 // the LAST maximum in iteration order wins
        sorted.push_back(val);
      });
      stable_sort_small(sorted.begin(), sorted.end(), [](double a, double b) { return a > b; });   // Collections.reverseOrder()
      if (!sorted.empty()) {
        double sum = 0.0;
        for (double v : sorted) sum += v;                              // sumUpList :750-756
        if (sum >= 0.2) {                                              // :232
          std::vector<double>& afi = f->afi;
          for (size_t k = 0; k < sorted.size(); ++k) {
            if (afi.size() > k) afi[k] = afi[k] + sorted[k];
            else if (afi.empty()) afi.insert(afi.end(), sorted.begin(), sorted.end());   // addAll: entries 1.. get added again
            else afi.push_back(sorted[k]);
          }
          f->totals.num_crosslinked_clusters++;
          for (uint64_t m = cl.mask51 & ((1ull << 51) - 1); m; m &= m - 1) {
            f->totals.allele_positions[__builtin_ctzll(m)]++;
            f->totals.num_allele_positions++;
          }
        }
      }
      for (double v : sorted) fraction += v;                           // :258-260
      row.best_pos = best_pos;
      row.best_value = best;
      if (best_pos > 0) {
        const int32_t* bc = mm.get(best_pos);
        row.best_count = bc ? (uint32_t)*bc : 0;
        row.has_ccr = 1;                                               // CCR row / FASTA are written (:262)
      }
    }
    row.fraction = fraction;
  }
  f->totals.n_allele_frequency = f->afi.size();
  return PS_OK;
}

// running totals; allele_frequency_information receives min(max, n_allele_frequency) RAW sums (the reference divides
// them by numCrosslinkedClusters when it writes <out>.sitefrequency, :529-536)
int ps_flush_totals_get(const ps_flush* f, ps_flush_totals* out, double* allele_frequency_information, uint64_t max) {
  if (!f || !out) return PS_ERR_INVALID_ARG;
  *out = f->totals;
  out->n_allele_frequency = f->afi.size();
  if (allele_frequency_information)
    for (uint64_t k = 0; k < std::min<uint64_t>(max, f->afi.size()); ++k) allele_frequency_information[k] = f->afi[k];
  return PS_OK;
}

}  // extern "C"
