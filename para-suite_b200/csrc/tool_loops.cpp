// The two tool loops over FILES, streaming and across several GPUs of one process (the entry points a JVM binds when
// it replaces ErrorProfiling.inferErrorProfile and PileupClusters.calculateReadPileups: Main.java:595-597, :634-636).
//
//   error profile : batches of records go round-robin to the contexts, every context accumulates its share, the int64
//                   accumulator vectors (< 10 KB) are summed on the host at the end (all outputs are commutative integer
//                   sums; Java's int wrap-around is applied after the sum).
//   T>C pileup    : the file is taken in WINDOWS of records (PileupClusters.java:137 streams; a window is a shard in
//                   time).  The carry-in of a window -- Java's (tempClusterChr, tempClusterEnd) at its first record -- is
//                   the maximum (contig, end) over all earlier windows, known on the host from the decoded records before
//                   the window is launched, so windows go to the contexts round-robin and run while the next one is
//                   being decoded.  The reads at the head of a window that continue the cluster left open by the window
//                   before come back as a "head partial" and are folded into that cluster here (halo merge: counts,
//                   masks, strand state, T>C sites by position with their first-insertion keys, coverage re-read from the
//                   summed dense coverage of the boundary cluster).
//   clust files   : the merged closed clusters of every window go, with the window's records, through the flush
//                   arithmetic and the native writers (flush.cpp, clust_writer.cpp).
// No compute happens on the host: boundaries, masks, counts and sites all come from the kernels.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "internal.h"

extern "C" int open_for_ctx(ps_ctx* ctx, const char* bam_path, uint64_t max_batch, ps_bam** B);
int stage_batch(ps_ctx* ctx, const ps_read_batch* hb, bool with_qual, StagedBatch** out);

struct ps_multi {
  std::vector<ps_ctx*> ctx;
  bool owns = true;
  ps_packed_fasta* fasta = nullptr;
  std::string err;
  std::mutex err_mu;          // the window loop's sink thread reports errors too
};

static int multi_fail(ps_multi* m, int st, const std::string& msg) {
  if (m) {
    std::lock_guard<std::mutex> g(m->err_mu);
    m->err = msg;
  }
  return st;
}

namespace {

// max over the kept records of a host batch of (contig + 1) << 32 | alignment end (what pl_maxkey_kernel computes)
uint64_t host_max_key(const ps_read_batch& b, const std::vector<uint64_t>& contig_off) {
  uint64_t best = 0, coff = (!b.uniform_ncigar && b.tile_cigar_off) ? b.tile_cigar_off[0] : 0;
  size_t ci = 0;
  for (uint64_t r = 0; r < b.n_reads; ++r) {
    const uint32_t meta = b.meta[r], flags = PS_META_FLAGS(meta), ncig = PS_META_NCIGAR(meta);
    const uint32_t* cig = b.cigar + coff;
    coff += ncig;
    if (flags & (PS_RF_UNMAPPED | PS_RF_POS_ZERO | PS_RF_CIGAR_OVERFLOW)) continue;
    bool hasI = false, hasD = false, hasN = false;
    uint64_t R = 0;
    for (uint32_t e = 0; e < ncig; ++e) {
      const uint32_t op = cig[e] & 15u;
      hasI |= op == 1u; hasD |= op == 2u; hasN |= op == 3u;
      if ((0x18Du >> op) & 1u) R += cig[e] >> 4;
    }
    if ((hasI || hasD) && hasN) continue;
    const uint64_t g = b.ref_start[r];
    if (g >= contig_off.back()) continue;
    if (!(g >= contig_off[ci] && g < contig_off[ci + 1]))
      ci = (size_t)(std::upper_bound(contig_off.begin(), contig_off.end(), g) - contig_off.begin()) - 1;
    const uint64_t end = g - contig_off[ci] + R;          // start + R - 1 with start = g - off + 1
    const uint64_t key = ((uint64_t)(ci + 1) << 32) | (uint32_t)end;
    best = std::max(best, key);
  }
  return best;
}

struct OpenCluster {            // the cluster left open behind the windows taken so far
  bool have = false;
  ps_cluster c{};
  std::vector<ps_site> sites;
  int32_t cov_pos0 = 0;
  std::vector<uint32_t> cov;    // dense baseCoveredMap: cov[k] = coverage at cov_pos0 + k
};

// halo merge: the head partial of the next window (counts of the reads that continue the open cluster) into it
void merge_head(OpenCluster& o, const ps_cluster& hp, const std::vector<ps_site>& hs, int32_t hp_pos0,
                const std::vector<uint32_t>& hcov, uint64_t& double_stranded) {
  o.c.num_reads += hp.num_reads;
  o.c.num_t2c += hp.num_t2c;
  o.c.end = std::max(o.c.end, hp.end);
  o.c.mask51 |= hp.mask51;
  o.c.minus_after_first += hp.minus_after_first;
  if (!o.c.first_reverse) double_stranded += hp.minus_after_first;      // doubleStranded++ per minus member (:494-498)
  o.c.combined_strand = o.c.first_reverse ? 1 : (o.c.minus_after_first ? 2 : 0);
  // coverage: a site seen on one side of the cut is also covered by the other side's reads
  if (!hcov.empty()) {
    if (o.cov.empty()) { o.cov = hcov; o.cov_pos0 = hp_pos0; }
    else {
      const int64_t lo = std::min<int64_t>(o.cov_pos0, hp_pos0);
      const int64_t hi = std::max<int64_t>((int64_t)o.cov_pos0 + (int64_t)o.cov.size(), (int64_t)hp_pos0 + (int64_t)hcov.size());
      std::vector<uint32_t> sum((size_t)(hi - lo), 0);
      for (size_t k = 0; k < o.cov.size(); ++k) sum[(size_t)(o.cov_pos0 - lo) + k] += o.cov[k];
      for (size_t k = 0; k < hcov.size(); ++k) sum[(size_t)(hp_pos0 - lo) + k] += hcov[k];
      o.cov.swap(sum);
      o.cov_pos0 = (int32_t)lo;
    }
  }
  // sites: union by position; counts add, the first-insertion key is the smaller one
  std::vector<ps_site> all = o.sites;
  all.insert(all.end(), hs.begin(), hs.end());
  std::stable_sort(all.begin(), all.end(), [](const ps_site& a, const ps_site& b) { return a.pos < b.pos; });
  std::vector<ps_site> merged;
  for (const ps_site& s : all) {
    if (!merged.empty() && merged.back().pos == s.pos) {
      merged.back().t2c += s.t2c;
      merged.back().order_key = std::min(merged.back().order_key, s.order_key);
    } else merged.push_back(s);
  }
  for (ps_site& s : merged) {
    const int64_t k = (int64_t)s.pos - o.cov_pos0;
    s.cov = (k >= 0 && k < (int64_t)o.cov.size()) ? o.cov[(size_t)k] : 0u;
  }
  o.sites.swap(merged);
}

// what one window hands to its consumer: the clusters that closed with it (site_begin / site_end index `sites`), and the
// first read of the cluster still open behind it
struct WindowOut {
  ps_read_batch host{};        // copy of the window's batch descriptor (the arrays live in the batcher's slab)
  const ps_read_batch* batch;
  uint64_t first_ordinal;
  std::vector<ps_cluster> closed;
  std::vector<ps_site> sites;
  bool has_open;
  uint64_t open_first_read;
};
using WindowSink = std::function<int(const WindowOut&)>;

struct InFlight {
  ps_ctx* ctx = nullptr;
  ps_pileup* h = nullptr;
  ps_read_batch host{};       // the window's records (slab of the batcher: valid until the third-next ps_bam_next)
  uint64_t offset = 0;
  bool live = false;
};

int sync_fail(ps_multi* m, ps_ctx* c, int st) { return multi_fail(m, st, ps_last_error(c)); }

// The windowed pileup over a file.  counters / open: the run's totals and the cluster left open at the end.
int pileup_windows(ps_multi* m, const char* bam_path, const ps_pileup_opts* opts, uint64_t window_reads, const WindowSink& sink,
                   ps_pileup_counters* counters, OpenCluster* open_out, ps_fault* fault) {
  if (m->ctx.empty()) return multi_fail(m, PS_ERR_STATE, "no context");
  ps_ctx* c0 = m->ctx[0];
  ps_bam* B = nullptr;
  int st = open_for_ctx(c0, bam_path, window_reads, &B);
  if (st) return sync_fail(m, c0, st);
  const std::vector<uint64_t>& contig_off = c0->contig_off;
  const uint32_t first_id = opts ? opts->first_running_id : 1u;
  uint64_t carry_key = 0;
  if (opts && opts->carry_valid) carry_key = ((uint64_t)(opts->carry_contig + 1) << 32) | (uint32_t)opts->carry_cluster_end;
  memset(counters, 0, sizeof *counters);
  OpenCluster open;
  uint64_t created = 0, offset = 0, window = 0;
  InFlight fl[2];
  // The sink (the host writer of `clust`: cluster sequences, rows, files) runs on a thread of its own, one window at a
  // time and in file order, while this thread decodes the next window: the batcher keeps three slabs, so the records
  // of the window in the sink's hands stay valid until its successor has been handed over (which joins it first).
  std::thread sink_thread;
  int sink_rc = PS_OK;
  auto join_sink = [&]() -> int {
    if (sink_thread.joinable()) sink_thread.join();
    const int rc = sink_rc;
    sink_rc = PS_OK;
    return rc;
  };

  // completes window `f`: fetches its records, merges the boundary, hands the closed clusters to the sink
  auto complete = [&](InFlight& f) -> int {
    f.live = false;
    ps_ctx* c = f.ctx;
    int rc = ps_pileup_wait(f.h);
    if (rc != PS_OK) {
      if (rc == PS_ERR_REFERENCE_WOULD_THROW || rc == PS_ERR_UNSUPPORTED) {
        ps_pileup_fault(f.h, fault);
        fault->read_ordinal += f.offset;
      }
      multi_fail(m, rc, ps_last_error(c));
      ps_pileup_close(f.h);
      return rc;
    }
    ps_pileup_counters ctr;
    ps_pileup_counters_get(f.h, &ctr);
    counters->num_reads_processed += ctr.num_reads_processed;
    counters->skipped_due_indel += ctr.skipped_due_indel;
    counters->double_stranded += ctr.double_stranded;
    auto out_p = std::make_shared<WindowOut>();
    WindowOut& out = *out_p;
    out.host = f.host;
    out.batch = &out.host;
    out.first_ordinal = f.offset;
    // head partial -> the open cluster
    {
      ps_cluster hp;
      std::vector<ps_site> hs(1 << 12);
      int k = ps_pileup_head_partial(f.h, &hp, hs.data(), hs.size());
      if (k == PS_ERR_INVALID_ARG) { hs.resize(1 << 20); k = ps_pileup_head_partial(f.h, &hp, hs.data(), hs.size()); }
      if (k < 0) { ps_pileup_close(f.h); return multi_fail(m, k, "head partial"); }
      if (k > 0) {
        if (!open.have) { ps_pileup_close(f.h); return multi_fail(m, PS_ERR_STATE, "head partial without an open cluster"); }
        hs.resize(hp.site_end);
        for (ps_site& s : hs) s.order_key += f.offset << 6;
        int32_t p0 = 0;
        const int64_t ln = ps_pileup_boundary_coverage(f.h, 0, &p0, nullptr, 0);
        std::vector<uint32_t> cov((size_t)std::max<int64_t>(ln, 0));
        if (ln > 0) ps_pileup_boundary_coverage(f.h, 0, &p0, cov.data(), cov.size());
        merge_head(open, hp, hs, p0, cov, counters->double_stranded);
      }
    }
    const uint64_t n_closed = ctr.n_clusters, n_new = n_closed + (ctr.has_open_cluster ? 1 : 0);
    if (n_new && open.have) {            // the open cluster closes with this window's first new cluster
      ps_cluster c2 = open.c;
      c2.site_begin = 0;
      c2.site_end = open.sites.size();
      out.closed.push_back(c2);
      out.sites = open.sites;
      open = OpenCluster();
    }
    if (n_closed) {
      const size_t c_at = out.closed.size(), s_at = out.sites.size();
      out.closed.resize(c_at + n_closed);
      out.sites.resize(s_at + ctr.n_sites);
      const int64_t got = ps_pileup_next(f.h, 0, out.closed.data() + c_at, n_closed, out.sites.data() + s_at, ctr.n_sites);
      if (got != (int64_t)n_closed) { ps_pileup_close(f.h); return multi_fail(m, got < 0 ? (int)got : PS_ERR_STATE, "ps_pileup_next"); }
      for (size_t k = c_at; k < out.closed.size(); ++k) {
        out.closed[k].running_id += (uint32_t)created;
        out.closed[k].first_read += f.offset;
        out.closed[k].site_begin += s_at;
        out.closed[k].site_end += s_at;
      }
      for (size_t k = s_at; k < out.sites.size(); ++k) out.sites[k].order_key += f.offset << 6;
    }
    if (ctr.has_open_cluster) {
      std::vector<ps_site> os(1 << 12);
      int k = ps_pileup_open_cluster(f.h, &open.c, os.data(), os.size());
      if (k == PS_ERR_INVALID_ARG) { os.resize(1 << 20); k = ps_pileup_open_cluster(f.h, &open.c, os.data(), os.size()); }
      if (k <= 0) { ps_pileup_close(f.h); return multi_fail(m, k < 0 ? k : PS_ERR_STATE, "open cluster"); }
      os.resize(open.c.site_end);
      for (ps_site& s : os) s.order_key += f.offset << 6;
      open.sites.swap(os);
      open.c.first_read += f.offset;
      open.c.running_id += (uint32_t)created;
      int32_t p0 = 0;
      const int64_t ln = ps_pileup_boundary_coverage(f.h, 1, &p0, nullptr, 0);
      open.cov.assign((size_t)std::max<int64_t>(ln, 0), 0);
      if (ln > 0) ps_pileup_boundary_coverage(f.h, 1, &p0, open.cov.data(), open.cov.size());
      open.cov_pos0 = p0;
      open.have = true;
    }
    ps_pileup_close(f.h);
    created += n_new;
    out.has_open = open.have;
    out.open_first_read = open.have ? open.c.first_read : 0;
    counters->n_clusters += out.closed.size();
    counters->n_sites += out.sites.size();
    const int prev_rc = join_sink();                       // the window before this one has been written
    if (prev_rc != PS_OK) return prev_rc;
    sink_thread = std::thread([&sink, &sink_rc, out_p] { sink_rc = sink(*out_p); });
    return PS_OK;
  };

  for (;; ++window) {
    InFlight& f = fl[window & 1];
    // two windows in flight at most: the slab ps_bam_next is about to fill belongs to the window before last, whose
    // records the sink still needs -- complete it first (windows complete in file order)
    if (f.live) { st = complete(f); if (st) break; }
    ps_read_batch hb;
    const int k = ps_bam_next(B, &hb);
    if (k < 0) { st = multi_fail(m, k, ps_bam_error(B)); break; }
    if (k == 0) break;
    ps_ctx* c = m->ctx[window % m->ctx.size()];
    // a context takes one submitted call at a time (its scratch belongs to it): with a single context the window before
    // -- whose kernels ran while this one was being decoded -- completes before this one is launched
    InFlight& prev = fl[(window & 1) ^ 1];
    if (prev.live && prev.ctx == c) { st = complete(prev); if (st) break; }
    cudaSetDevice(c->device);
    // cluster ids are numbered from the window's own first cluster; the clusters opened by earlier windows are added
    // when the window completes (they are not known yet: those windows may still be running)
    ps_pileup_opts o{};
    o.first_running_id = first_id;
    if (carry_key) { o.carry_valid = 1; o.carry_contig = (uint32_t)(carry_key >> 32) - 1; o.carry_cluster_end = (int32_t)(uint32_t)carry_key; }
    f.ctx = c; f.host = hb; f.offset = offset;
    // upload without the quality bytes (the pileup never reads them: more than half of the records' bytes)
    StagedBatch* sb = nullptr;
    st = stage_batch(c, &hb, /*with_qual=*/false, &sb);
    if (st == PS_OK) {
      ps_read_batch view = hb;
      const DeviceBatch& v = sb->view;
      view.meta = v.meta; view.ref_start = v.ref_start; view.bases2 = v.bases2; view.qual = v.qual; view.cigar = v.cigar;
      view.tile_base_off = v.tile_base_off; view.tile_qual_off = v.tile_qual_off; view.tile_cigar_off = v.tile_cigar_off;
      view.tile_exc_off = v.tile_exc_off; view.exc = v.exc;
      st = ps_pileup_submit_device(c, &view, &o, nullptr, &f.h);
    }
    if (st != PS_OK) { sync_fail(m, c, st); break; }
    f.live = true;
    carry_key = std::max(carry_key, host_max_key(hb, contig_off));
    offset += hb.n_reads;
  }
  for (int i = 0; i < 2 && st == PS_OK; ++i) {
    InFlight& f = fl[(window + i) & 1];
    if (f.live) st = complete(f);
  }
  {
    const int rc = join_sink();
    if (st == PS_OK) st = rc;
  }
  for (InFlight& f : fl)
    if (f.live) { ps_pileup_close(f.h); f.live = false; }
  ps_bam_close(B);
  if (st != PS_OK) return st;
  counters->has_open_cluster = open.have ? 1 : 0;
  if (open_out) *open_out = std::move(open);
  return PS_OK;
}

}  // namespace

extern "C" {

// ---- several devices behind one handle --------------------------------------------------------------------------------
int ps_create_multi(ps_multi** out, const int* devices, int n) {
  if (!out) return PS_ERR_INVALID_ARG;
  *out = nullptr;
  std::vector<int> dev;
  if (devices && n > 0) dev.assign(devices, devices + n);
  else if (const char* e = getenv("PARASUITE_B200_DEVICES")) {      // "0,1,2"
    for (const char* p = e; *p;) {
      char* q = nullptr;
      const long v = strtol(p, &q, 10);
      if (q == p) break;
      dev.push_back((int)v);
      p = (*q == ',') ? q + 1 : q;
    }
  }
  if (dev.empty()) dev.push_back(0);
  ps_multi* m = new ps_multi();
  for (int d : dev) {
    ps_ctx* c = nullptr;
    const int st = ps_create(&c, d);
    if (st != PS_OK) {
      for (ps_ctx* x : m->ctx) ps_destroy(x);
      delete m;
      return st;
    }
    m->ctx.push_back(c);
  }
  *out = m;
  return PS_OK;
}

void ps_destroy_multi(ps_multi* m) {
  if (!m) return;
  if (m->owns)
    for (ps_ctx* c : m->ctx) ps_destroy(c);
  if (m->fasta) ps_fasta_free(m->fasta);
  delete m;
}

int ps_multi_device_count(const ps_multi* m) { return m ? (int)m->ctx.size() : 0; }
ps_ctx* ps_multi_context(ps_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[(size_t)i] : nullptr; }
const char* ps_multi_last_error(const ps_multi* m) { return m ? m->err.c_str() : "no object"; }

// FASTA (+ .fai) packed once on the host, resident in the HBM of every device
int ps_multi_load_fasta(ps_multi* m, const char* fasta_path) {
  if (!m || !fasta_path) return PS_ERR_INVALID_ARG;
  ps_packed_fasta* F = nullptr;
  int st = ps_fasta_pack(fasta_path, &F);
  if (st != PS_OK) {
    multi_fail(m, st, F ? ps_fasta_error(F) : "FASTA pack failed");
    ps_fasta_free(F);
    return st;
  }
  for (ps_ctx* c : m->ctx) {
    st = ps_reference_upload(c, ps_fasta_reference(F));
    if (st != PS_OK) { sync_fail(m, c, st); ps_fasta_free(F); return st; }
    if (c->fasta && !c->fasta_shared) ps_fasta_free(c->fasta);
    c->fasta = F;
    c->fasta_shared = true;
    c->fasta_path = fasta_path;
  }
  if (m->fasta) ps_fasta_free(m->fasta);
  m->fasta = F;
  return PS_OK;
}

// ErrorProfiling.inferErrorProfile's record loop (:104-409): batches round-robin over the devices, one host sum at the end
int ps_multi_profile_bam(ps_multi* m, const char* bam_path, const ps_profile_opts* opts, ps_profile_result* out) {
  if (!m || !bam_path || !opts || !out) return PS_ERR_INVALID_ARG;
  if (m->ctx.empty()) return multi_fail(m, PS_ERR_STATE, "no context");
  ps_bam* B = nullptr;
  uint64_t batch_reads = 1ull << 22;
  if (const char* e = getenv("PARASUITE_B200_BATCH_READS")) { const long long v = atoll(e); if (v > 0) batch_reads = (uint64_t)v; }
  int st = open_for_ctx(m->ctx[0], bam_path, batch_reads, &B);
  if (st) return sync_fail(m, m->ctx[0], st);
  ps_profile_opts o = *opts;
  o.emit_t2c_masks = 0;
  for (ps_ctx* c : m->ctx) {
    st = ps_profile_begin(c, &o);
    if (st) { sync_fail(m, c, st); break; }
  }
  uint64_t ordinal = 0, batch = 0;
  cudaEvent_t pending[2] = {nullptr, nullptr};        // upload of the batch that owns the slab about to be reused
  int pending_dev[2] = {0, 0};
  ps_read_batch hb;
  while (st == PS_OK) {
    if (pending[batch & 1]) { cudaSetDevice(pending_dev[batch & 1]); cudaEventSynchronize(pending[batch & 1]); pending[batch & 1] = nullptr; }
    const int k = ps_bam_next(B, &hb);
    if (k < 0) { st = multi_fail(m, k, ps_bam_error(B)); break; }
    if (k == 0) break;
    ps_ctx* c = m->ctx[batch % m->ctx.size()];
    c->reads_seen = ordinal;                           // fault ordinals are positions in the file, not in the context's share
    const int slot = c->staged_next;
    st = ps_profile_batch(c, &hb);
    if (st) { sync_fail(m, c, st); break; }
    pending[batch & 1] = c->staged_done[slot];
    pending_dev[batch & 1] = c->device;
    ordinal += hb.n_reads;
    ++batch;
  }
  // every context's accumulator, summed; the earliest fault wins
  const size_t n_acc = ps_profile_acc_len(o.max_read_length, o.infer_qualities);
  std::vector<int64_t> sum(n_acc, 0), part(n_acc);
  ps_fault first_fault{};
  int first_fault_st = PS_OK;
  std::string first_fault_msg;
  for (ps_ctx* c : m->ctx) {
    if (!c->profile_open) continue;
    ps_profile_result r;
    memset(&r, 0, sizeof r);
    r.wide = part.data();
    const int e = ps_profile_end(c, &r);
    if (e == PS_ERR_REFERENCE_WOULD_THROW || e == PS_ERR_UNSUPPORTED) {
      if (first_fault_st == PS_OK || r.fault.read_ordinal < first_fault.read_ordinal) {
        first_fault = r.fault; first_fault_st = e; first_fault_msg = ps_last_error(c);
      }
    } else if (e != PS_OK && st == PS_OK) st = sync_fail(m, c, e);
    else if (e == PS_OK)
      for (size_t k = 0; k < n_acc; ++k) sum[k] += part[k];
  }
  ps_bam_close(B);
  if (st != PS_OK) return st;
  out->fault.code = 0;
  out->fault.read_ordinal = 0;
  if (first_fault_st != PS_OK) {
    out->fault = first_fault;
    return multi_fail(m, first_fault_st, first_fault_msg);
  }
  profile_fill_result(make_layout(o.max_read_length, o.infer_qualities ? 1 : 0), sum.data(), out);
  return PS_OK;
}

static uint64_t window_reads_default() {
  uint64_t w = 1ull << 23;
  if (const char* e = getenv("PARASUITE_B200_WINDOW_READS")) { const long long v = atoll(e); if (v > 0) w = (uint64_t)v; }
  return std::min<uint64_t>(w, 0xFFFFFF00ull);
}

// PileupClusters.calculateReadPileups' record loop (:62-500), streaming: the handle holds the merged records of all
// windows on the host (ps_pileup_next / ps_pileup_open_cluster / ps_pileup_counters_get as usual)
int ps_multi_pileup_bam(ps_multi* m, const char* bam_path, const ps_pileup_opts* opts, ps_pileup** out) {
  if (!m || !bam_path || !out) return PS_ERR_INVALID_ARG;
  *out = nullptr;
  std::vector<ps_cluster> clusters;
  std::vector<ps_site> sites;
  ps_pileup_counters ctr;
  OpenCluster open;
  ps_fault fault{};
  const int st = pileup_windows(m, bam_path, opts, window_reads_default(), [&](const WindowOut& w) {
    const size_t s_at = sites.size();
    for (ps_cluster c : w.closed) { c.site_begin += s_at; c.site_end += s_at; clusters.push_back(c); }
    sites.insert(sites.end(), w.sites.begin(), w.sites.end());
    return PS_OK;
  }, &ctr, &open, &fault);
  if (st != PS_OK) {
    if (st == PS_ERR_REFERENCE_WOULD_THROW || st == PS_ERR_UNSUPPORTED) {      // a handle that only carries the fault
      ps_pileup_counters none{};
      ps_pileup* H = pileup_host_handle(m->ctx[0], {}, {}, false, ps_cluster{}, {}, 0, {}, none);
      *out = H;
      pileup_set_fault(H, fault);
    }
    return st;
  }
  if (open.have) { open.c.site_begin = 0; open.c.site_end = open.sites.size(); }
  *out = pileup_host_handle(m->ctx[0], std::move(clusters), std::move(sites), open.have, open.c, std::move(open.sites),
                            open.cov_pos0, std::move(open.cov), ctr);
  return PS_OK;
}

// The whole `clust` tool from files (PileupClusters.calculateReadPileups :62-584): <out>, <out>.ccr.fasta, <out>.ccr.tsv,
// <out>.report, <bam>.sitefrequency.tsv, <bam>.sitepositions.tsv.  snp_vcf may be NULL (no SNP filter).
int ps_multi_clust_bam(ps_multi* m, const char* bam_path, const char* out_path, const char* snp_vcf, uint32_t min_read_coverage,
                       ps_pileup_counters* counters_out, ps_fault* fault_out) {
  if (!m || !bam_path || !out_path) return PS_ERR_INVALID_ARG;
  if (m->ctx.empty() || !m->ctx[0]->fasta) return multi_fail(m, PS_ERR_STATE, "load the reference first (ps_multi_load_fasta)");
  ps_ctx* c0 = m->ctx[0];
  std::vector<const char*> names;
  for (uint32_t i = 0; i < c0->ref.n_contigs; ++i) names.push_back(ps_fasta_contig_name(c0->fasta, i));
  ps_flush* fl = nullptr;
  int st = ps_flush_create(&fl, min_read_coverage, (uint32_t)names.size(), names.data());
  if (st != PS_OK) return multi_fail(m, st, "ps_flush_create");
  if (snp_vcf && *snp_vcf) {
    st = ps_flush_load_vcf(fl, snp_vcf);
    if (st != PS_OK) { multi_fail(m, st, ps_flush_error(fl)); ps_flush_destroy(fl); return st; }
  }
  ps_clust_writer* w = nullptr;
  st = ps_clust_writer_open(&w, fl, c0->fasta_path.c_str(), out_path, bam_path);
  if (st != PS_OK) { multi_fail(m, st, ps_clust_writer_error(w)); ps_clust_writer_close(w); ps_flush_destroy(fl); return st; }
  ps_pileup_counters ctr;
  ps_fault fault{};
  st = pileup_windows(m, bam_path, nullptr, window_reads_default(), [&](const WindowOut& o) {
    const int rc = ps_clust_writer_feed(w, o.batch, o.first_ordinal, o.closed.data(), o.closed.size(), o.sites.data(),
                                        o.has_open ? 1 : 0, o.open_first_read);
    if (rc != PS_OK) {
      multi_fail(m, rc, ps_clust_writer_error(w));
      if (rc == PS_ERR_REFERENCE_WOULD_THROW) ps_clust_writer_fault(w, &fault);
    }
    return rc;
  }, &ctr, nullptr, &fault);
  if (st == PS_OK) {
    st = ps_clust_writer_finish(w, &ctr);
    if (st != PS_OK) multi_fail(m, st, ps_clust_writer_error(w));
  }
  if (counters_out) *counters_out = ctr;
  if (fault_out) *fault_out = fault;
  ps_clust_writer_close(w);
  ps_flush_destroy(fl);
  return st;
}

// ---- the single-context forms ------------------------------------------------------------------------------------------
static void borrow(ps_multi& m, ps_ctx* ctx) { m.ctx.assign(1, ctx); m.owns = false; }

int ps_pileup_bam(ps_ctx* ctx, const char* bam_path, const ps_pileup_opts* opts, ps_pileup** out) {
  if (!ctx || !bam_path || !out) return set_error(ctx, PS_ERR_INVALID_ARG, "NULL argument");
  ps_multi m;
  borrow(m, ctx);
  const int st = ps_multi_pileup_bam(&m, bam_path, opts, out);
  if (st != PS_OK && !m.err.empty()) set_error(ctx, st, m.err);
  return st;
}

// The whole `error` tool from files (ErrorProfiling.inferErrorProfile, :96-621 without the plot): record loop on the
// GPU(s), the six output files natively.
int ps_multi_error_bam(ps_multi* m, const char* bam_path, const ps_profile_opts* opts, int32_t* counters_out, ps_fault* fault_out) {
  if (!m || !bam_path || !opts) return PS_ERR_INVALID_ARG;
  const uint32_t ml = opts->max_read_length;
  std::vector<int32_t> pc((size_t)ml * 16), qpm(16), qpmc(16), ctr(PS_PC_COUNT);
  std::vector<double> ins(ml), del(ml);
  std::vector<int64_t> qh(opts->infer_qualities ? (size_t)ml * 256 : 0);
  ps_profile_result r{};
  r.position_conversions = pc.data(); r.quality_per_mismatch = qpm.data(); r.quality_per_mismatch_counts = qpmc.data();
  r.insertions_per_pos = ins.data(); r.deletions_per_pos = del.data(); r.counters = ctr.data();
  r.quality_hist = opts->infer_qualities ? qh.data() : nullptr;
  int st = ps_multi_profile_bam(m, bam_path, opts, &r);
  if (fault_out) *fault_out = r.fault;
  if (st != PS_OK) return st;
  char err[512];
  st = ps_profile_write_files(&r, ml, opts->infer_qualities, bam_path, nullptr, err, sizeof err);
  if (st != PS_OK) return multi_fail(m, st, err);
  if (counters_out) memcpy(counters_out, ctr.data(), sizeof(int32_t) * PS_PC_COUNT);
  return PS_OK;
}

int ps_error_bam(ps_ctx* ctx, const char* bam_path, const ps_profile_opts* opts, int32_t* counters_out, ps_fault* fault_out) {
  if (!ctx || !bam_path || !opts) return set_error(ctx, PS_ERR_INVALID_ARG, "NULL argument");
  ps_multi m;
  borrow(m, ctx);
  const int st = ps_multi_error_bam(&m, bam_path, opts, counters_out, fault_out);
  if (st != PS_OK && !m.err.empty()) set_error(ctx, st, m.err);
  return st;
}

int ps_clust_bam(ps_ctx* ctx, const char* bam_path, const char* out_path, const char* snp_vcf, uint32_t min_read_coverage,
                 ps_pileup_counters* counters_out, ps_fault* fault_out) {
  if (!ctx || !bam_path || !out_path) return set_error(ctx, PS_ERR_INVALID_ARG, "NULL argument");
  ps_multi m;
  borrow(m, ctx);
  const int st = ps_multi_clust_bam(&m, bam_path, out_path, snp_vcf, min_read_coverage, counters_out, fault_out);
  if (st != PS_OK && !m.err.empty()) set_error(ctx, st, m.err);
  return st;
}

}  // extern "C"
