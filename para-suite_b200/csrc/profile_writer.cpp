// Output files of the `error` tool after the record loop (ErrorProfiling.java:410-591), from the arrays the profile
// kernels fill: <bam>.errorprofile, .errorprofile.vcf, .qualityPerMismatch, .indels, .indelprofile and, with -q,
// .qualities.  `.errorprofile` and `.indelprofile` are what the error-tolerant aligner consumes (`bwa parasuite -p/-g`,
// PARAsuiteMapping.java:69-72).  Host only; doubles are printed the way Java prints them.  The same arithmetic as
// parasuite_b200/profile_files.py (which the tests hold against the literal restatement in oracle/py_oracle.py).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "java_double.h"
#include "parasuite_b200.h"

namespace {

bool write_text(const std::string& path, const std::string& text, std::string& err) {
  FILE* f = fopen(path.c_str(), "w");
  if (!f) { err = "cannot create " + path; return false; }
  const bool ok = fwrite(text.data(), 1, text.size(), f) == text.size();
  if (fclose(f) != 0 || !ok) { err = "short write to " + path; return false; }
  return true;
}

double div(double a, double b) { return a / b; }      // IEEE: x/0 = +-Infinity, 0/0 = NaN, like Java

}  // namespace

extern "C" int ps_profile_write_files(const ps_profile_result* res, uint32_t max_read_length, uint32_t infer_qualities,
                                      const char* bam_path, double* averaged_t2c_epr, char* err, size_t err_cap) {
  if (err && err_cap) err[0] = 0;
  if (!res || !bam_path || !res->position_conversions || !res->quality_per_mismatch || !res->quality_per_mismatch_counts ||
      !res->insertions_per_pos || !res->deletions_per_pos || !res->counters || (infer_qualities && !res->quality_hist))
    return PS_ERR_INVALID_ARG;
  const uint32_t m = max_read_length;
  const int32_t* pc = res->position_conversions;        // [m][ref][read]
  static const char B[] = "ACGT";
  double tot[4][4] = {}, tot_base[4] = {};
  std::vector<int32_t> tot_pos(m, 0);                   // per-position totals wrap like the Java int field totalCountsPerPos
  for (uint32_t i = 0; i < m; ++i) {
    int64_t t = 0;
    for (int j = 0; j < 4; ++j)
      for (int k = 0; k < 4; ++k) { tot[j][k] += (double)pc[i * 16 + j * 4 + k]; t += pc[i * 16 + j * 4 + k]; }
    tot_pos[i] = (int32_t)(uint32_t)(uint64_t)t;
  }
  for (int j = 0; j < 4; ++j)
    for (int k = 0; k < 4; ++k) tot_base[j] += tot[j][k];
  std::string vcf, prof, qpm;
  for (int j = 0; j < 4; ++j) {
    for (int k = 0; k < 4; ++k) {
      vcf += std::string(1, B[j]) + "\t" + std::string(1, B[k]) + "\t" + java_double(tot[j][k]) + "\n";
      prof += java_double(div(tot[j][k], tot_base[j])) + "\t";
      qpm += java_double(div((double)res->quality_per_mismatch[j * 4 + k], (double)res->quality_per_mismatch_counts[j * 4 + k])) + "\t";
    }
    vcf += "\n";
    prof += "\n";
    qpm += "\n";
  }
  // averaged T>C errors per read over the leading positions with T>C > 0 (:532-542)
  {
    const double n_proc = (double)res->counters[PS_PC_NUM_READS_PROCESSED];
    double acc = 0.0, avg = 0.0;
    uint32_t j = 0;
    for (; j < m; ++j) {
      const double v = div((double)pc[j * 16 + 3 * 4 + 1], n_proc) * 100.0;
      if (!(v > 0)) break;
      acc += v;
    }
    avg = j < m ? acc / (double)(j + 1) : acc;
    if (averaged_t2c_epr) *averaged_t2c_epr = avg;
  }
  // indel rates per position (:553-589)
  std::string indels;
  double ia = 0.0, da = 0.0;
  uint32_t iz = 0, dz = 0;
  for (uint32_t i = 0; i < m; ++i) {
    const bool seen = tot_pos[i] != 0;
    const double ir = seen ? div(res->insertions_per_pos[i], (double)tot_pos[i]) : 0.0;
    const double dr = seen ? div(res->deletions_per_pos[i], (double)tot_pos[i]) : 0.0;
    indels += java_double(ir) + "\t" + java_double(dr) + "\n";
    if (ir > 0) ia += ir; else ++iz;       // sequential sums in position order, like the Java loop
    if (dr > 0) da += dr; else ++dz;
  }
  if (iz == m && dz == m) ia = da = 0.0;
  else { ia = div(ia, (double)(m - iz)); da = div(da, (double)(m - dz)); }
  const std::string indelprofile = java_double(ia) + "\t" + java_double(da);
  std::string qualities;
  if (infer_qualities) {   // the mean is exact; the SD is summed over the histogram, not over Java's linked list
    for (uint32_t i = 0; i < m; ++i) {
      const int64_t* h = res->quality_hist + (size_t)i * 256;
      double n = 0, sum = 0;
      for (int q = 0; q < 256; ++q) { const double v = q < 128 ? q : q - 256; n += (double)h[q]; sum += (double)h[q] * v; }
      const double mean = div(sum, n);
      double var = 0;
      for (int q = 0; q < 256; ++q) { const double v = q < 128 ? q : q - 256; var += (double)h[q] * (v - mean) * (v - mean); }
      qualities += java_double(mean) + "\t" + java_double(std::sqrt(div(var, n))) + "\n";
    }
  }
  const std::string base(bam_path);
  std::string e;
  const bool ok = write_text(base + ".errorprofile", prof, e) && write_text(base + ".errorprofile.vcf", vcf, e) &&
                  write_text(base + ".qualityPerMismatch", qpm, e) && write_text(base + ".indels", indels, e) &&
                  write_text(base + ".indelprofile", indelprofile, e) && write_text(base + ".qualities", qualities, e);
  if (!ok) {
    if (err && err_cap) snprintf(err, err_cap, "%s", e.c_str());
    return PS_ERR_IO;
  }
  return PS_OK;
}
