// `comb` tool, host side: transcript hits lifted to genomic coordinates and merged with the genomic hits.
// Replaces CombineGenomeTranscript.combine / printReadsToBamFile
// (/root/reference/src/src/utils/postprocessing/CombineGenomeTranscript.java:36-666, Read.java) -- the step directly
// upstream of both kernels: it writes the `aMbNcM` cigars they must tolerate.  Serial string / cigar
// surgery on BAM records, I/O bound: no GPU work here.  The quirks of the Java are kept on purpose:
//   * exon starts and exon ends are sorted as STRINGS, each list on its own (:172-175 Arrays.sort(String[]))
//   * a hit whose walk stops at "indel + splice junction" (:294, :431, :447) keeps the start and the partial cigar it had
//     collected so far (plus strand) -- it is still written
//   * minus-strand transcripts: the strand flag flips and the BASES are reverse-complemented, the qualities stay (:650-659)
//   * a read is written only when all its hits lift to the same start (:603-611); the record used is the last primary
//     hit (Read.primaryIndex), MAPQ becomes 10
//   * contig = "chr" + field 3 of the transcript name; absent from the genomic header -> dropped; "MT" becomes "M" after
//     that test (:621-631)
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "parasuite_b200.h"

namespace {

struct Fail { int st; std::string msg; };

std::vector<std::string> java_split(const std::string& s, char sep) {   // String.split: trailing empty strings removed
  std::vector<std::string> out;
  size_t a = 0;
  for (;;) {
    const size_t b = s.find(sep, a);
    out.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
    if (b == std::string::npos) break;
    a = b + 1;
  }
  while (!out.empty() && out.back().empty()) out.pop_back();
  if (out.empty() && s.empty()) out.push_back("");
  return out;
}

int32_t java_parse_int(const std::string& s) {   // Integer.parseInt; NumberFormatException kills the tool
  size_t k = (!s.empty() && (s[0] == '+' || s[0] == '-')) ? 1 : 0;
  if (k == s.size()) throw Fail{PS_ERR_REFERENCE_WOULD_THROW, "NumberFormatException: \"" + s + "\""};
  int64_t v = 0;
  for (size_t i = k; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') throw Fail{PS_ERR_REFERENCE_WOULD_THROW, "NumberFormatException: \"" + s + "\""};
    v = v * 10 + (s[i] - '0');
    if (v > (int64_t)1 << 31) throw Fail{PS_ERR_REFERENCE_WOULD_THROW, "NumberFormatException: \"" + s + "\""};
  }
  if (s[0] == '-') v = -v;
  if (v > INT32_MAX) throw Fail{PS_ERR_REFERENCE_WOULD_THROW, "NumberFormatException: \"" + s + "\""};
  return (int32_t)v;
}

struct Lift { int32_t start = -1; std::string cigar; uint32_t missed = 0; };

// CombineGenomeTranscript.java:146-474 for one hit
Lift lift_hit(const std::string& ref_name, int32_t aln_start, int32_t aln_end, int32_t read_len, const std::string& cigar) {
  const std::vector<std::string> f = java_split(ref_name, '|');
  if (f.size() < 6) throw Fail{PS_ERR_REFERENCE_WOULD_THROW, "ArrayIndexOutOfBoundsException: transcript name without six |-separated fields: " + ref_name};
  std::vector<std::string> es = java_split(f[3], ';'), ee = java_split(f[4], ';');
  std::sort(es.begin(), es.end());
  std::sort(ee.begin(), ee.end());
  const std::string& strand = f[5];
  const int n = (int)es.size();
  auto st = [&](int i) { return java_parse_int(es[i]); };
  auto en = [&](int i) {
    if (i >= (int)ee.size()) throw Fail{PS_ERR_REFERENCE_WOULD_THROW, "ArrayIndexOutOfBoundsException: fewer exon ends than starts: " + ref_name};
    return java_parse_int(ee[i]);
  };
  const bool has_indel = cigar.find('D') != std::string::npos || cigar.find('I') != std::string::npos;
  Lift L;
  int32_t passed = 0;
  if (strand == "1") {                                             // :219-391
    for (int i = 0; i < n; ++i) {
      const int32_t tmp = passed;
      passed += en(i) - st(i) + 1;
      if (aln_start <= passed && L.start == -1) L.start = st(i) + (aln_start - tmp) - 1;
      if (aln_end <= passed) {
        if (L.start >= st(i)) L.cigar = cigar;
        else L.cigar += std::to_string(aln_end - tmp) + "M";
        break;
      } else if (L.start != -1) {
        if (has_indel) { L.missed++; break; }
        if (L.start >= st(i)) L.cigar += std::to_string(en(i) - L.start + 1) + "M";
        else L.cigar += std::to_string(en(i) - st(i) + 1) + "M";
        if (i < n - 1) {
          const int32_t intron = st(i + 1) - en(i) - 1;
          if (intron <= 0) break;
          L.cigar += std::to_string(intron) + "N";
        } else break;
      }
    }
  } else if (strand == "-1") {                                     // :392-473
    int32_t new_end = -1;
    for (int i = n - 1; i >= 0; --i) {
      const int32_t tmp = passed;
      passed += en(i) - st(i) + 1;
      if (aln_start <= passed && new_end == -1) new_end = en(i) - (aln_start - tmp) + 1;
      if (aln_end <= passed) {
        if (new_end <= en(i)) { L.cigar = cigar; L.start = new_end - read_len + 1; }
        else {
          if (has_indel) { L.missed++; break; }
          L.cigar = std::to_string(aln_end - tmp) + "M" + L.cigar;
          L.start = en(i) - (aln_end - tmp) + 1;
        }
        break;
      } else if (new_end != -1) {
        if (has_indel) { L.missed++; break; }
        if (new_end < en(i)) L.cigar = std::to_string(new_end - st(i) + 1) + "M" + L.cigar;
        else L.cigar = std::to_string(en(i) - st(i) + 1) + "M" + L.cigar;
        if (i >= 1) {
          const int32_t intron = st(i) - en(i - 1) - 1;
          if (intron <= 0) break;
          L.cigar = std::to_string(intron) + "N" + L.cigar;
        } else break;
      }
    }
  }
  return L;
}

// ---- BAM in memory ---------------------------------------------------------------------------------------------
uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
int32_t rdi32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
void put16(std::vector<uint8_t>& b, size_t at, uint16_t v) { memcpy(b.data() + at, &v, 2); }
void put32(std::vector<uint8_t>& b, size_t at, uint32_t v) { memcpy(b.data() + at, &v, 4); }

struct Bam {
  std::string text;                       // header text
  std::vector<std::string> names;         // reference names
  std::vector<uint32_t> lens;
  std::vector<uint8_t> data;              // inflated file
  std::vector<size_t> rec;                // offset of every record's block_size field
  std::string sort_order() const {        // SAMFileHeader.getSortOrder: @HD SO, "unsorted" when absent
    if (text.compare(0, 3, "@HD") != 0) return "unsorted";
    const size_t eol = text.find('\n');
    const std::string line = text.substr(0, eol);
    const size_t so = line.find("\tSO:");
    if (so == std::string::npos) return "unsorted";
    const size_t e = line.find('\t', so + 4);
    return line.substr(so + 4, e == std::string::npos ? std::string::npos : e - so - 4);
  }
};

void load_bam(const char* path, Bam& B) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) throw Fail{PS_ERR_IO, std::string("cannot open ") + path};
  struct stat sb;
  fstat(fd, &sb);
  const size_t n = (size_t)sb.st_size;
  const uint8_t* p = n ? (const uint8_t*)mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
  close(fd);
  if (n && p == MAP_FAILED) throw Fail{PS_ERR_IO, std::string("cannot map ") + path};
  struct Unmap { const uint8_t* p; size_t n; ~Unmap() { if (p) munmap((void*)p, n); } } um{p, n};
  struct Blk { size_t coff; uint32_t clen, isize; size_t at; };
  std::vector<Blk> blocks;
  size_t o = 0, total = 0;
  while (o < n) {
    if (o + 18 > n || p[o] != 0x1f || p[o + 1] != 0x8b || p[o + 2] != 8 || !(p[o + 3] & 4))
      throw Fail{PS_ERR_FORMAT, std::string(path) + ": not a BGZF file"};
    const uint32_t xlen = rd16(p + o + 10);
    uint32_t bsize = 0;
    size_t x = o + 12;
    const size_t xend = x + xlen;
    if (xend > n) throw Fail{PS_ERR_FORMAT, std::string(path) + ": truncated BGZF block"};
    while (x + 4 <= xend) {
      const uint32_t slen = rd16(p + x + 2);
      if (x + 4 + (size_t)slen > xend) throw Fail{PS_ERR_FORMAT, std::string(path) + ": bad BGZF extra field"};
      if (p[x] == 'B' && p[x + 1] == 'C' && slen == 2) bsize = (uint32_t)rd16(p + x + 4) + 1;
      x += 4 + slen;
    }
    if (bsize < xlen + 20 || o + bsize > n) throw Fail{PS_ERR_FORMAT, std::string(path) + ": truncated BGZF block"};
    const uint32_t isize = rd32(p + o + bsize - 4);
    if (isize) { blocks.push_back({o + 12 + xlen, bsize - xlen - 20, isize, total}); total += isize; }
    o += bsize;
  }
  B.data.resize(total);
  const int threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  std::vector<int> bad(threads, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      if (inflateInit2(&zs, -15) != Z_OK) { bad[t] = 1; return; }
      for (size_t k = blocks.size() * t / threads; k < blocks.size() * (t + 1) / threads; ++k) {
        inflateReset(&zs);
        zs.next_in = const_cast<Bytef*>(p + blocks[k].coff);
        zs.avail_in = blocks[k].clen;
        zs.next_out = B.data.data() + blocks[k].at;
        zs.avail_out = blocks[k].isize;
        if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.avail_out != 0) { bad[t] = 1; break; }
      }
      inflateEnd(&zs);
    });
  for (auto& th : pool) th.join();
  for (int b : bad)
    if (b) throw Fail{PS_ERR_FORMAT, std::string(path) + ": corrupt BGZF block"};
  const std::vector<uint8_t>& d = B.data;
  if (d.size() < 12 || memcmp(d.data(), "BAM\1", 4) != 0) throw Fail{PS_ERR_FORMAT, std::string(path) + ": not a BAM file"};
  const uint32_t l_text = rd32(d.data() + 4);
  if (d.size() < 12 + (size_t)l_text) throw Fail{PS_ERR_FORMAT, std::string(path) + ": truncated BAM header"};
  B.text.assign((const char*)d.data() + 8, l_text);
  while (!B.text.empty() && B.text.back() == '\0') B.text.pop_back();
  size_t q = 8 + l_text;
  const uint32_t n_ref = rd32(d.data() + q);
  q += 4;
  for (uint32_t r = 0; r < n_ref; ++r) {
    if (d.size() < q + 4) throw Fail{PS_ERR_FORMAT, std::string(path) + ": truncated BAM header"};
    const uint32_t l_name = rd32(d.data() + q);
    if (d.size() < q + 8 + (size_t)l_name) throw Fail{PS_ERR_FORMAT, std::string(path) + ": truncated BAM header"};
    B.names.emplace_back((const char*)d.data() + q + 4, l_name ? l_name - 1 : 0);
    B.lens.push_back(rd32(d.data() + q + 4 + l_name));
    q += 8 + (size_t)l_name;
  }
  while (q + 4 <= d.size()) {
    const uint32_t bs = rd32(d.data() + q);
    if (bs < 32 || q + 4 + (size_t)bs > d.size()) throw Fail{PS_ERR_FORMAT, std::string(path) + ": truncated BAM record"};
    B.rec.push_back(q);
    q += 4 + (size_t)bs;
  }
}

// fields of a raw record (offsets from the block_size field)
struct RecView {
  const uint8_t* p;
  uint32_t block_size() const { return rd32(p); }
  int32_t ref_id() const { return rdi32(p + 4); }
  int32_t pos0() const { return rdi32(p + 8); }
  uint32_t l_name() const { return p[12]; }
  uint32_t mapq() const { return p[13]; }
  uint32_t n_cigar() const { return rd16(p + 16); }
  uint32_t flag() const { return rd16(p + 18); }
  uint32_t l_seq() const { return rd32(p + 20); }
  int32_t next_ref() const { return rdi32(p + 24); }
  int32_t next_pos() const { return rdi32(p + 28); }
  int32_t tlen() const { return rdi32(p + 32); }
  const char* name() const { return (const char*)p + 36; }
  const uint8_t* cigar() const { return p + 36 + l_name(); }
  const uint8_t* seq() const { return cigar() + 4 * (size_t)n_cigar(); }
  const uint8_t* qual() const { return seq() + (l_seq() + 1) / 2; }
  const uint8_t* tags() const { return qual() + l_seq(); }
  const uint8_t* end() const { return p + 4 + block_size(); }
  bool sane() const { return 36 + (size_t)l_name() + 4 * (size_t)n_cigar() + (l_seq() + 1) / 2 + l_seq() <= 4 + (size_t)block_size(); }
};

std::string cigar_string(const RecView& r) {       // SAMRecord.getCigarString: "*" without elements
  if (r.n_cigar() == 0) return "*";
  std::string s;
  for (uint32_t k = 0; k < r.n_cigar(); ++k) {
    const uint32_t c = rd32(r.cigar() + 4 * k);
    s += std::to_string(c >> 4);
    s += "MIDNSHP=X????????"[c & 15];
  }
  return s;
}
uint32_t cigar_ref_len(const RecView& r) {
  uint32_t R = 0;
  for (uint32_t k = 0; k < r.n_cigar(); ++k) {
    const uint32_t c = rd32(r.cigar() + 4 * k), op = c & 15;
    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) R += c >> 4;
  }
  return R;
}
std::vector<uint32_t> parse_cigar(const std::string& s) {      // TextCigarCodec.decode on the strings built above
  std::vector<uint32_t> ops;
  if (s.empty() || s == "*") return ops;
  uint64_t num = 0;
  bool have = false;
  for (char ch : s) {
    if (ch >= '0' && ch <= '9') { num = num * 10 + (ch - '0'); have = true; continue; }
    const char* tbl = "MIDNSHP=X";
    const char* at = strchr(tbl, ch);
    if (!at || !have || num >= (1u << 28)) throw Fail{PS_ERR_REFERENCE_WOULD_THROW, "IllegalArgumentException: malformed cigar " + s};
    ops.push_back((uint32_t)(num << 4) | (uint32_t)(at - tbl));
    num = 0;
    have = false;
  }
  if (have) throw Fail{PS_ERR_REFERENCE_WOULD_THROW, "IllegalArgumentException: malformed cigar " + s};
  return ops;
}

int reg2bin(int64_t beg, int64_t end) {             // SAM spec 5.3
  --end;
  if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

// the output: records as (sort key, bytes)
struct OutRec {
  int32_t ref_id, pos0;
  uint32_t flag, mapq;
  int32_t next_ref, next_pos, tlen;
  size_t off, len;      // in the arena
  size_t name_off;      // of the read name inside the arena
  uint64_t seqno;
};

struct BgzfWriter {
  FILE* f = nullptr;
  std::vector<uint8_t> buf;
  ~BgzfWriter() { if (f) fclose(f); }
  void flush_block(const uint8_t* p, size_t n) {
    uint8_t out[0x10000 + 64];
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, 5, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw Fail{PS_ERR_IO, "deflateInit2 failed"};
    zs.next_in = const_cast<Bytef*>(p);
    zs.avail_in = (uInt)n;
    zs.next_out = out + 18;
    zs.avail_out = sizeof out - 18 - 8;
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); throw Fail{PS_ERR_IO, "deflate failed"}; }
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    const uint8_t head[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
    memcpy(out, head, 12);
    out[12] = 'B'; out[13] = 'C'; out[14] = 2; out[15] = 0;
    const uint16_t bsize = (uint16_t)(clen + 25);
    memcpy(out + 16, &bsize, 2);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), p, (uInt)n), isize = (uint32_t)n;
    memcpy(out + 18 + clen, &crc, 4);
    memcpy(out + 22 + clen, &isize, 4);
    if (fwrite(out, 1, clen + 26, f) != clen + 26) throw Fail{PS_ERR_IO, "short write"};
  }
  void write(const void* p, size_t n) {
    const uint8_t* q = (const uint8_t*)p;
    while (n) {
      const size_t k = std::min(n, (size_t)0xff00 - buf.size());
      buf.insert(buf.end(), q, q + k);
      q += k;
      n -= k;
      if (buf.size() == 0xff00) { flush_block(buf.data(), buf.size()); buf.clear(); }
    }
  }
  void finish() {
    if (!buf.empty()) { flush_block(buf.data(), buf.size()); buf.clear(); }
    flush_block(nullptr, 0);      // the EOF marker block
    if (fclose(f) != 0) { f = nullptr; throw Fail{PS_ERR_IO, "close failed"}; }
    f = nullptr;
  }
};

void comb(const char* genomic, const char* transcript, const char* out_path, ps_comb_stats* S) {
  Bam G, T;
  load_bam(genomic, G);
  load_bam(transcript, T);
  if (T.sort_order() != "queryname")       // :90-99 (the Java logs the error and exits)
    throw Fail{PS_ERR_UNSORTED, std::string("BAM file ") + transcript + " is not sorted. Please provide a sorted BAM-file as input alignment file."};
  std::unordered_map<std::string, int32_t> gidx;
  for (size_t i = 0; i < G.names.size(); ++i) gidx.emplace(G.names[i], (int32_t)i);     // first wins on duplicates
  std::vector<uint8_t> arena;
  std::vector<OutRec> out;
  uint64_t seqno = 0;
  auto add_raw = [&](const uint8_t* p, size_t len) {
    RecView r{p};
    OutRec o{r.ref_id(), r.pos0(), r.flag(), r.mapq(), r.next_ref(), r.next_pos(), r.tlen(), arena.size(), len, arena.size() + 36, seqno++};
    arena.insert(arena.end(), p, p + len);
    out.push_back(o);
  };
  for (size_t q : G.rec) {                 // :53-60 every genomic record goes through unchanged
    RecView r{G.data.data() + q};
    if (!r.sane()) throw Fail{PS_ERR_FORMAT, std::string(genomic) + ": malformed BAM record"};
    add_raw(r.p, 4 + (size_t)r.block_size());
  }
  S->genomic_records = G.rec.size();
  S->mapped_reads = G.rec.size();
  S->transcript_records = T.rec.size();

  struct Hit { int32_t start; std::string cigar; size_t rec; };
  std::vector<Hit> group;
  size_t primary = 0;
  auto flush = [&]() {                     // printReadsToBamFile :598-666 for the one read of the group
    if (group.empty()) return;
    bool same = true;
    for (const Hit& h : group) same &= h.start == group[0].start;
    const Hit& h = group[primary];
    if (same) {
      RecView r{T.data.data() + h.rec};
      const std::vector<std::string> f = java_split(T.names[r.ref_id()], '|');
      auto it = gidx.find("chr" + f[2]);
      if (it != gidx.end()) {
        std::string chrom = f[2] == "MT" ? "M" : f[2];
        auto it2 = gidx.find("chr" + chrom);
        const int32_t ref_id = it2 == gidx.end() ? -1 : it2->second;
        const std::vector<uint32_t> ops = parse_cigar(h.cigar);
        if (ops.size() > 0xFFFF) throw Fail{PS_ERR_UNSUPPORTED, "lifted cigar with more than 65535 elements"};
        uint32_t R = 0;
        for (uint32_t c : ops) { const uint32_t op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) R += c >> 4; }
        const bool minus = f[5] == "-1";
        uint32_t flag = r.flag();
        if (minus) flag ^= 16u;
        // mate reference: by NAME in the genomic header (SAMRecord.setHeader re-resolves the names)
        int32_t next_ref = -1;
        if (r.next_ref() >= 0 && (size_t)r.next_ref() < T.names.size()) {
          auto itn = gidx.find(T.names[r.next_ref()]);
          if (itn != gidx.end()) next_ref = itn->second;
        }
        const size_t l_seq = r.l_seq(), seq_bytes = (l_seq + 1) / 2;
        const size_t tags_len = (size_t)(r.end() - r.tags());
        const size_t len = 36 + r.l_name() + 4 * ops.size() + seq_bytes + l_seq + tags_len;
        const size_t off = arena.size();
        arena.resize(off + len);
        uint8_t* d = arena.data() + off;
        const int32_t pos0 = h.start - 1;
        const uint32_t bs = (uint32_t)(len - 4);
        memcpy(d, &bs, 4);
        memcpy(d + 4, &ref_id, 4);
        memcpy(d + 8, &pos0, 4);
        d[12] = (uint8_t)r.l_name();
        d[13] = 10;                                                       // setMappingQuality(10)
        const uint16_t bin = (uint16_t)reg2bin(pos0 < 0 ? 0 : pos0, (pos0 < 0 ? 0 : pos0) + (R ? R : 1));
        memcpy(d + 14, &bin, 2);
        const uint16_t nc = (uint16_t)ops.size(), fl = (uint16_t)flag;
        memcpy(d + 16, &nc, 2);
        memcpy(d + 18, &fl, 2);
        const uint32_t ls = (uint32_t)l_seq;
        memcpy(d + 20, &ls, 4);
        const int32_t np = r.next_pos(), tl = r.tlen();
        memcpy(d + 24, &next_ref, 4);
        memcpy(d + 28, &np, 4);
        memcpy(d + 32, &tl, 4);
        memcpy(d + 36, r.name(), r.l_name());
        uint8_t* w = d + 36 + r.l_name();
        for (uint32_t c : ops) { memcpy(w, &c, 4); w += 4; }
        if (!minus) memcpy(w, r.seq(), seq_bytes);
        else {   // SequenceUtil.reverseComplement on the ASCII bases: A<->T, C<->G, every other symbol stays
          static const uint8_t comp[16] = {0, 8, 4, 3, 2, 5, 6, 7, 1, 9, 10, 11, 12, 13, 14, 15};
          memset(w, 0, seq_bytes);
          for (size_t k = 0; k < l_seq; ++k) {
            const size_t src = l_seq - 1 - k;
            const uint8_t nib = (r.seq()[src >> 1] >> ((~src & 1) * 4)) & 15;
            w[k >> 1] |= (uint8_t)(comp[nib] << ((~k & 1) * 4));
          }
        }
        w += seq_bytes;
        memcpy(w, r.qual(), l_seq);
        w += l_seq;
        memcpy(w, r.tags(), tags_len);
        out.push_back(OutRec{ref_id, pos0, flag, 10u, next_ref, np, tl, off, len, off + 36, seqno++});
        S->lifted_records++;
        S->mapped_reads++;
      }
    }
    group.clear();
    primary = 0;
  };
  std::string name_tmp;
  for (size_t q : T.rec) {
    RecView r{T.data.data() + q};
    if (!r.sane()) throw Fail{PS_ERR_FORMAT, std::string(transcript) + ": malformed BAM record"};
    if (r.ref_id() < 0) continue;                                         // :105 reference name "*"
    if ((size_t)r.ref_id() >= T.names.size()) throw Fail{PS_ERR_FORMAT, std::string(transcript) + ": reference index out of range"};
    const std::string name(r.name());
    if (name != name_tmp) { flush(); name_tmp = name; }                    // :137-144
    const int32_t start = r.pos0() + 1;
    const int32_t end = (r.flag() & 4u) ? 0 : start + (int32_t)cigar_ref_len(r) - 1;
    const Lift L = lift_hit(T.names[r.ref_id()], start, end, (int32_t)r.l_seq(), cigar_string(r));
    S->missed_transcript_alignments += L.missed;
    if (L.start == -1) continue;                                          // :476
    if (L.cigar.find('N') != std::string::npos) S->spliced_reads++;
    if (!(r.flag() & 0x100u)) primary = group.size();                     // Read.setPrimaryIndex(genesHitted.size())
    group.push_back(Hit{L.start, L.cigar, q});
  }
  flush();

  if (G.sort_order() == "coordinate") {    // the writer was opened with presorted = false: it sorts by the header's order
    const uint8_t* A = arena.data();
    std::stable_sort(out.begin(), out.end(), [&](const OutRec& a, const OutRec& b) {
      // SAMRecordCoordinateComparator: reference index (no reference last), start, strand, name, flags, MAPQ, mate
      const uint32_t ra = a.ref_id < 0 ? 0x7FFFFFFFu : (uint32_t)a.ref_id, rb = b.ref_id < 0 ? 0x7FFFFFFFu : (uint32_t)b.ref_id;
      if (ra != rb) return ra < rb;
      if (a.pos0 != b.pos0) return a.pos0 < b.pos0;
      const uint32_t sa = (a.flag >> 4) & 1u, sb = (b.flag >> 4) & 1u;
      if (sa != sb) return sa < sb;
      const int c = strcmp((const char*)A + a.name_off, (const char*)A + b.name_off);
      if (c) return c < 0;
      if (a.flag != b.flag) return a.flag < b.flag;
      if (a.mapq != b.mapq) return a.mapq < b.mapq;
      if (a.next_ref != b.next_ref) return a.next_ref < b.next_ref;
      if (a.next_pos != b.next_pos) return a.next_pos < b.next_pos;
      if (a.tlen != b.tlen) return a.tlen < b.tlen;
      return a.seqno < b.seqno;
    });
  } else if (G.sort_order() == "queryname") {
    const uint8_t* A = arena.data();
    std::stable_sort(out.begin(), out.end(), [&](const OutRec& a, const OutRec& b) {
      return strcmp((const char*)A + a.name_off, (const char*)A + b.name_off) < 0;
    });
  }
  BgzfWriter W;
  W.f = fopen(out_path, "wb");
  if (!W.f) throw Fail{PS_ERR_IO, std::string("cannot create ") + out_path};
  {  // the genomic header, text and reference list
    std::vector<uint8_t> h(12 + G.text.size());
    memcpy(h.data(), "BAM\1", 4);
    put32(h, 4, (uint32_t)G.text.size());
    memcpy(h.data() + 8, G.text.data(), G.text.size());
    put32(h, 8 + G.text.size(), (uint32_t)G.names.size());
    W.write(h.data(), h.size());
    for (size_t i = 0; i < G.names.size(); ++i) {
      const uint32_t ln = (uint32_t)G.names[i].size() + 1;
      W.write(&ln, 4);
      W.write(G.names[i].c_str(), ln);
      W.write(&G.lens[i], 4);
    }
  }
  for (const OutRec& o : out) W.write(arena.data() + o.off, o.len);
  W.finish();
}

}  // namespace

extern "C" {

int ps_liftover_hit(const char* transcript_name, int32_t aln_start, int32_t aln_end, int32_t read_len, const char* cigar,
                    int32_t* new_start, char* new_cigar, size_t new_cigar_cap, uint32_t* missed) {
  if (!transcript_name || !cigar || !new_start || !new_cigar || new_cigar_cap == 0) return PS_ERR_INVALID_ARG;
  try {
    const Lift L = lift_hit(transcript_name, aln_start, aln_end, read_len, cigar);
    if (L.cigar.size() + 1 > new_cigar_cap) return PS_ERR_INVALID_ARG;
    *new_start = L.start;
    memcpy(new_cigar, L.cigar.c_str(), L.cigar.size() + 1);
    if (missed) *missed = L.missed;
    return PS_OK;
  } catch (const Fail& f) {
    snprintf(new_cigar, new_cigar_cap, "%s", f.msg.c_str());
    return f.st;
  }
}

int ps_comb_bam(const char* genomic_bam, const char* transcript_bam, const char* out_bam, ps_comb_stats* stats, char* err,
                size_t err_cap) {
  ps_comb_stats local;
  memset(&local, 0, sizeof local);
  if (err && err_cap) err[0] = 0;
  if (!genomic_bam || !transcript_bam || !out_bam) return PS_ERR_INVALID_ARG;
  int st = PS_OK;
  try {
    comb(genomic_bam, transcript_bam, out_bam, &local);
  } catch (const Fail& f) {
    if (err && err_cap) snprintf(err, err_cap, "%s", f.msg.c_str());
    st = f.st;
  } catch (const std::bad_alloc&) {
    if (err && err_cap) snprintf(err, err_cap, "out of host memory");
    st = PS_ERR_OOM;
  }
  if (stats) *stats = local;
  return st;
}

}  // extern "C"
