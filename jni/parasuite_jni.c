/*
 * parasuite_jni.c -- thin JNI shim over libparasuite_b200.so (no logic of its own).
 *
 * Binds the two native classes a maintainer adds to PARA-suite (INTEGRATION.md shows the Java side):
 *
 *   utils.errorprofile.NativeErrorProfile   replaces the loop ErrorProfiling.java:146-409
 *   utils.pileupclusters.NativePileup       replaces the loop PileupClusters.java:137-500 (+ :585-673)
 *
 * Build (needs a JDK; none exists in the build image of this repo, so it is compiled only when jni.h is found):
 *   cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *      jni/parasuite_jni.c -Lpara-suite_b200/lib -lparasuite_b200 -o libparasuite_jni.so
 *
 * Error convention: every non-zero status becomes a java.lang.RuntimeException carrying ps_last_error();
 * PS_ERR_UNSORTED keeps the reference's message (ErrorProfiling.java:128-131) so the Java caller can log it and exit.
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "parasuite_b200.h"

static void throw_status(JNIEnv* env, ps_ctx* ctx, int status) {
  jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
  const char* msg = ctx ? ps_last_error(ctx) : ps_strerror(status);
  if (!msg || !*msg) msg = ps_strerror(status);
  if (cls) (*env)->ThrowNew(env, cls, msg);
}

static void throw_message(JNIEnv* env, const char* msg, int status) {
  jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
  if (!msg || !*msg) msg = ps_strerror(status);
  if (cls) (*env)->ThrowNew(env, cls, msg);
}

static void throw_illegal(JNIEnv* env, const char* msg) {
  jclass cls = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
  if (cls) (*env)->ThrowNew(env, cls, msg);
}

/* ---- context ------------------------------------------------------------------------------------------- */

/* static native long create(int device); */
JNIEXPORT jlong JNICALL Java_utils_errorprofile_NativeErrorProfile_create(JNIEnv* env, jclass c, jint device) {
  (void)c;
  ps_ctx* ctx = NULL;
  int st = ps_create(&ctx, (int)device);
  if (st != PS_OK) { throw_status(env, NULL, st); return 0; }
  return (jlong)(intptr_t)ctx;
}

/* static native void destroy(long ctx); */
JNIEXPORT void JNICALL Java_utils_errorprofile_NativeErrorProfile_destroy(JNIEnv* env, jclass c, jlong ctx) {
  (void)env; (void)c;
  ps_destroy((ps_ctx*)(intptr_t)ctx);
}

/* static native void loadReference(long ctx, String fasta);   -- new IndexedFastaSequenceFile(fasta), :109-110 */
JNIEXPORT void JNICALL Java_utils_errorprofile_NativeErrorProfile_loadReference(JNIEnv* env, jclass c, jlong h, jstring fasta) {
  (void)c;
  ps_ctx* ctx = (ps_ctx*)(intptr_t)h;
  const char* path = (*env)->GetStringUTFChars(env, fasta, NULL);
  if (!path) return;
  int st = ps_reference_load_fasta(ctx, path);
  (*env)->ReleaseStringUTFChars(env, fasta, path);
  if (st != PS_OK) throw_status(env, ctx, st);
}

/* ---- error profile --------------------------------------------------------------------------------------
 * static native void profileBam(long ctx, String bam, int maxReadLength, boolean inferQualities,
 *     int[] positionConversions (maxLen*16, [i][ref][read]), int[] qualityPerMismatch (16),
 *     int[] qualityPerMismatchCounts (16), double[] insertionsPerPos (maxLen), double[] deletionsPerPos (maxLen),
 *     long[] qualityHist (maxLen*256 or null), int[] counters (8: numReadsProcessed, unmapped, duplicates,
 *     startZero, indelRead, skippedReads, longerIndels, totalBasesChecked));
 * The arrays are the Java fields themselves (ErrorProfiling.java:41-55); the loop's post-processing (:410-621)
 * runs unchanged on them. */
static void profile_bam_common(JNIEnv* env, int multi, jlong h, jstring bam, jint maxReadLength, jboolean inferQualities,
                               jintArray positionConversions, jintArray qualityPerMismatch, jintArray qualityPerMismatchCounts,
                               jdoubleArray insertionsPerPos, jdoubleArray deletionsPerPos, jlongArray qualityHist,
                               jintArray counters) {
  ps_ctx* ctx = multi ? NULL : (ps_ctx*)(intptr_t)h;
  ps_multi* mc = multi ? (ps_multi*)(intptr_t)h : NULL;
  const char* path = (*env)->GetStringUTFChars(env, bam, NULL);
  if (!path) return;
  ps_profile_opts opts;
  memset(&opts, 0, sizeof opts);
  opts.max_read_length = (uint32_t)maxReadLength;
  opts.infer_qualities = inferQualities ? 1u : 0u;
  /* the library writes 16*maxLen, 16, 16, maxLen, maxLen, 8 and 256*maxLen elements: a shorter Java array would mean a
   * write past its end, i.e. a corrupted JVM heap */
  const jsize m = (jsize)maxReadLength;
  if (maxReadLength <= 0 || !positionConversions || !qualityPerMismatch || !qualityPerMismatchCounts || !insertionsPerPos ||
      !deletionsPerPos || !counters || (*env)->GetArrayLength(env, positionConversions) < 16 * m ||
      (*env)->GetArrayLength(env, qualityPerMismatch) < 16 || (*env)->GetArrayLength(env, qualityPerMismatchCounts) < 16 ||
      (*env)->GetArrayLength(env, insertionsPerPos) < m || (*env)->GetArrayLength(env, deletionsPerPos) < m ||
      (*env)->GetArrayLength(env, counters) < PS_PC_COUNT ||
      (inferQualities && (!qualityHist || (*env)->GetArrayLength(env, qualityHist) < 256 * m))) {
    (*env)->ReleaseStringUTFChars(env, bam, path);
    throw_illegal(env, "profileBam: a result array is missing or shorter than maxReadLength implies");
    return;
  }
  ps_profile_result r;
  memset(&r, 0, sizeof r);
  /* plain Get/Release (not Critical): the call runs for a while and must not block the collector */
  r.position_conversions = (int32_t*)(*env)->GetIntArrayElements(env, positionConversions, NULL);
  r.quality_per_mismatch = (int32_t*)(*env)->GetIntArrayElements(env, qualityPerMismatch, NULL);
  r.quality_per_mismatch_counts = (int32_t*)(*env)->GetIntArrayElements(env, qualityPerMismatchCounts, NULL);
  r.insertions_per_pos = (double*)(*env)->GetDoubleArrayElements(env, insertionsPerPos, NULL);
  r.deletions_per_pos = (double*)(*env)->GetDoubleArrayElements(env, deletionsPerPos, NULL);
  r.counters = (int32_t*)(*env)->GetIntArrayElements(env, counters, NULL);
  r.quality_hist = (inferQualities && qualityHist) ? (int64_t*)(*env)->GetLongArrayElements(env, qualityHist, NULL) : NULL;
  /* Get*ArrayElements returns NULL (with an OutOfMemoryError pending) when the VM cannot pin or copy */
  const int got_all = r.position_conversions && r.quality_per_mismatch && r.quality_per_mismatch_counts && r.insertions_per_pos &&
                      r.deletions_per_pos && r.counters && (!(inferQualities && qualityHist) || r.quality_hist);
  int st = !got_all ? PS_OK : (multi ? ps_multi_profile_bam(mc, path, &opts, &r) : ps_profile_bam(ctx, path, &opts, &r));
  if (r.position_conversions) (*env)->ReleaseIntArrayElements(env, positionConversions, (jint*)r.position_conversions, 0);
  if (r.quality_per_mismatch) (*env)->ReleaseIntArrayElements(env, qualityPerMismatch, (jint*)r.quality_per_mismatch, 0);
  if (r.quality_per_mismatch_counts) (*env)->ReleaseIntArrayElements(env, qualityPerMismatchCounts, (jint*)r.quality_per_mismatch_counts, 0);
  if (r.insertions_per_pos) (*env)->ReleaseDoubleArrayElements(env, insertionsPerPos, (jdouble*)r.insertions_per_pos, 0);
  if (r.deletions_per_pos) (*env)->ReleaseDoubleArrayElements(env, deletionsPerPos, (jdouble*)r.deletions_per_pos, 0);
  if (r.counters) (*env)->ReleaseIntArrayElements(env, counters, (jint*)r.counters, 0);
  if (r.quality_hist) (*env)->ReleaseLongArrayElements(env, qualityHist, (jlong*)r.quality_hist, 0);
  (*env)->ReleaseStringUTFChars(env, bam, path);
  if (got_all && st != PS_OK) {                                /* !got_all: the VM's own exception is pending */
    if (multi) throw_message(env, ps_multi_last_error(mc), st);
    else throw_status(env, ctx, st);
  }
}

JNIEXPORT void JNICALL Java_utils_errorprofile_NativeErrorProfile_profileBam(
    JNIEnv* env, jclass c, jlong h, jstring bam, jint maxReadLength, jboolean inferQualities, jintArray positionConversions,
    jintArray qualityPerMismatch, jintArray qualityPerMismatchCounts, jdoubleArray insertionsPerPos,
    jdoubleArray deletionsPerPos, jlongArray qualityHist, jintArray counters) {
  (void)c;
  profile_bam_common(env, 0, h, bam, maxReadLength, inferQualities, positionConversions, qualityPerMismatch,
                     qualityPerMismatchCounts, insertionsPerPos, deletionsPerPos, qualityHist, counters);
}

/* ---- several GPUs in one JVM (ps_create_multi): the same calls on a multi handle ---------------------------------
 * static native long createMulti(int[] devices);   -- null or empty: PARASUITE_B200_DEVICES ("0,1,2"), else device 0
 * static native void destroyMulti(long multi);
 * static native void loadReferenceMulti(long multi, String fasta);
 * static native void profileBamMulti(long multi, String bam, ... as profileBam ...); */
JNIEXPORT jlong JNICALL Java_utils_errorprofile_NativeErrorProfile_createMulti(JNIEnv* env, jclass c, jintArray devices) {
  (void)c;
  ps_multi* m = NULL;
  jint* d = devices ? (*env)->GetIntArrayElements(env, devices, NULL) : NULL;
  const jsize n = devices ? (*env)->GetArrayLength(env, devices) : 0;
  int st = ps_create_multi(&m, (const int*)d, (int)n);
  if (d) (*env)->ReleaseIntArrayElements(env, devices, d, 0);
  if (st != PS_OK) { throw_status(env, NULL, st); return 0; }
  return (jlong)(intptr_t)m;
}

JNIEXPORT void JNICALL Java_utils_errorprofile_NativeErrorProfile_destroyMulti(JNIEnv* env, jclass c, jlong m) {
  (void)env; (void)c;
  ps_destroy_multi((ps_multi*)(intptr_t)m);
}

JNIEXPORT void JNICALL Java_utils_errorprofile_NativeErrorProfile_loadReferenceMulti(JNIEnv* env, jclass c, jlong h, jstring fasta) {
  (void)c;
  ps_multi* m = (ps_multi*)(intptr_t)h;
  const char* path = (*env)->GetStringUTFChars(env, fasta, NULL);
  if (!path) return;
  int st = ps_multi_load_fasta(m, path);
  (*env)->ReleaseStringUTFChars(env, fasta, path);
  if (st != PS_OK) throw_message(env, ps_multi_last_error(m), st);
}

JNIEXPORT void JNICALL Java_utils_errorprofile_NativeErrorProfile_profileBamMulti(
    JNIEnv* env, jclass c, jlong h, jstring bam, jint maxReadLength, jboolean inferQualities, jintArray positionConversions,
    jintArray qualityPerMismatch, jintArray qualityPerMismatchCounts, jdoubleArray insertionsPerPos,
    jdoubleArray deletionsPerPos, jlongArray qualityHist, jintArray counters) {
  (void)c;
  profile_bam_common(env, 1, h, bam, maxReadLength, inferQualities, positionConversions, qualityPerMismatch,
                     qualityPerMismatchCounts, insertionsPerPos, deletionsPerPos, qualityHist, counters);
}

/* ---- T>C pileup ------------------------------------------------------------------------------------------
 * static native long pileupBam(long ctx, String bam);        -> handle
 * static native long[] counters(long handle);                -> numReadsProcessed, skippedDueIndel, doubleStranded,
 *                                                               nClusters, nSites, hasOpenCluster
 * static native int nextClusters(long handle, long first, long[] cluster64 (8 per cluster), int maxClusters,
 *                                long[] site64 (3 per site), int maxSites);   -> clusters copied
 * static native void close(long handle);
 * Cluster and site records cross as their raw 64-bit words (ps_cluster = 8 words, ps_site = 3 words); the Java
 * side unpacks them with shifts. */
JNIEXPORT jlong JNICALL Java_utils_pileupclusters_NativePileup_pileupBam(JNIEnv* env, jclass c, jlong h, jstring bam) {
  (void)c;
  ps_ctx* ctx = (ps_ctx*)(intptr_t)h;
  const char* path = (*env)->GetStringUTFChars(env, bam, NULL);
  if (!path) return 0;
  ps_pileup_opts opts;
  memset(&opts, 0, sizeof opts);
  opts.first_running_id = 1; /* runningID = 1 (PileupClusters.java:133) */
  ps_pileup* out = NULL;
  int st = ps_pileup_bam(ctx, path, &opts, &out);
  (*env)->ReleaseStringUTFChars(env, bam, path);
  if (st != PS_OK) {
    if (out) ps_pileup_close(out);
    throw_status(env, ctx, st);
    return 0;
  }
  return (jlong)(intptr_t)out;
}

JNIEXPORT jlongArray JNICALL Java_utils_pileupclusters_NativePileup_counters(JNIEnv* env, jclass c, jlong handle) {
  (void)c;
  ps_pileup_counters ctr;
  memset(&ctr, 0, sizeof ctr);
  ps_pileup_counters_get((ps_pileup*)(intptr_t)handle, &ctr);
  jlong v[6] = {(jlong)ctr.num_reads_processed, (jlong)ctr.skipped_due_indel, (jlong)ctr.double_stranded,
                (jlong)ctr.n_clusters, (jlong)ctr.n_sites, (jlong)ctr.has_open_cluster};
  jlongArray a = (*env)->NewLongArray(env, 6);
  if (a) (*env)->SetLongArrayRegion(env, a, 0, 6, v);
  return a;
}

JNIEXPORT jint JNICALL Java_utils_pileupclusters_NativePileup_nextClusters(JNIEnv* env, jclass c, jlong handle, jlong first,
                                                                             jlongArray cluster64, jint maxClusters,
                                                                             jlongArray site64, jint maxSites) {
  (void)c;
  /* ps_cluster = 8 and ps_site = 3 64-bit words: the arrays must hold maxClusters / maxSites records */
  if (!cluster64 || maxClusters < 0 || maxSites < 0 || (maxSites > 0 && !site64) ||
      (jlong)(*env)->GetArrayLength(env, cluster64) < 8 * (jlong)maxClusters ||
      (site64 && (jlong)(*env)->GetArrayLength(env, site64) < 3 * (jlong)maxSites)) {
    throw_illegal(env, "nextClusters: cluster64 / site64 are shorter than maxClusters * 8 / maxSites * 3");
    return 0;
  }
  jlong* cl = (*env)->GetLongArrayElements(env, cluster64, NULL);
  jlong* si = site64 ? (*env)->GetLongArrayElements(env, site64, NULL) : NULL;
  int64_t n = 0;
  const int got_all = cl && (si || !site64);
  if (got_all)
    n = ps_pileup_next((ps_pileup*)(intptr_t)handle, (uint64_t)first, (ps_cluster*)cl, (uint64_t)maxClusters, (ps_site*)si,
                       si ? (uint64_t)maxSites : 0);
  if (cl) (*env)->ReleaseLongArrayElements(env, cluster64, cl, 0);
  if (si) (*env)->ReleaseLongArrayElements(env, site64, si, 0);
  if (!got_all) return 0;                                       /* the VM's OutOfMemoryError is pending */
  if (n < 0) { throw_status(env, NULL, (int)n); return 0; }
  return (jint)n;
}

/* static native long[] clustBam(long ctx, boolean multi, String bam, String out, String snpVcf, int minReadCoverage);
 * The whole body of PileupClusters.calculateReadPileups (:62-584): record loop on the GPU(s) in windows, flush and the six
 * output files natively.  Returns numReadsProcessed, skippedDueIndel, doubleStranded, nClusters, nSites, hasOpenCluster.
 * `multi`: ctx is a createMulti handle. */
JNIEXPORT jlongArray JNICALL Java_utils_pileupclusters_NativePileup_clustBam(JNIEnv* env, jclass c, jlong h, jboolean multi,
                                                                              jstring bam, jstring out, jstring snpVcf,
                                                                              jint minReadCoverage) {
  (void)c;
  const char* pb = (*env)->GetStringUTFChars(env, bam, NULL);
  const char* po = pb ? (*env)->GetStringUTFChars(env, out, NULL) : NULL;
  const char* pv = (po && snpVcf) ? (*env)->GetStringUTFChars(env, snpVcf, NULL) : NULL;
  jlongArray a = NULL;
  if (pb && po && (pv || !snpVcf)) {
    ps_pileup_counters ctr;
    ps_fault fault;
    memset(&ctr, 0, sizeof ctr);
    memset(&fault, 0, sizeof fault);
    int st = multi ? ps_multi_clust_bam((ps_multi*)(intptr_t)h, pb, po, pv, (uint32_t)minReadCoverage, &ctr, &fault)
                   : ps_clust_bam((ps_ctx*)(intptr_t)h, pb, po, pv, (uint32_t)minReadCoverage, &ctr, &fault);
    if (st != PS_OK) {
      if (multi) throw_message(env, ps_multi_last_error((ps_multi*)(intptr_t)h), st);
      else throw_status(env, (ps_ctx*)(intptr_t)h, st);
    } else {
      jlong v[6] = {(jlong)ctr.num_reads_processed, (jlong)ctr.skipped_due_indel, (jlong)ctr.double_stranded,
                    (jlong)ctr.n_clusters, (jlong)ctr.n_sites, (jlong)ctr.has_open_cluster};
      a = (*env)->NewLongArray(env, 6);
      if (a) (*env)->SetLongArrayRegion(env, a, 0, 6, v);
    }
  }
  if (pv) (*env)->ReleaseStringUTFChars(env, snpVcf, pv);
  if (po) (*env)->ReleaseStringUTFChars(env, out, po);
  if (pb) (*env)->ReleaseStringUTFChars(env, bam, pb);
  return a;
}

JNIEXPORT void JNICALL Java_utils_pileupclusters_NativePileup_close(JNIEnv* env, jclass c, jlong handle) {
  (void)env; (void)c;
  ps_pileup_close((ps_pileup*)(intptr_t)handle);
}

/* package utils.postprocessing;  class NativeCombine { static native long[] combine(String genomicBam, String transcriptBam,
 * String combinedBam); }  -- the whole body of CombineGenomeTranscript.combine (:36-596).  Returns mappedReads,
 * splicedReads, missedTranscriptAlignments, liftedRecords.  Host only: no context needed. */
JNIEXPORT jlongArray JNICALL Java_utils_postprocessing_NativeCombine_combine(JNIEnv* env, jclass c, jstring genomic,
                                                                             jstring transcript, jstring combined) {
  (void)c;
  if (!genomic || !transcript || !combined) { throw_illegal(env, "combine: file names must not be null"); return NULL; }
  const char* pg = (*env)->GetStringUTFChars(env, genomic, NULL);
  const char* pt = pg ? (*env)->GetStringUTFChars(env, transcript, NULL) : NULL;
  const char* po = pt ? (*env)->GetStringUTFChars(env, combined, NULL) : NULL;
  jlongArray a = NULL;
  if (pg && pt && po) {
    ps_comb_stats s;
    char err[512];
    const int st = ps_comb_bam(pg, pt, po, &s, err, sizeof err);
    if (st != PS_OK) throw_message(env, err, st);
    else {
      jlong v[4] = {(jlong)s.mapped_reads, (jlong)s.spliced_reads, (jlong)s.missed_transcript_alignments, (jlong)s.lifted_records};
      a = (*env)->NewLongArray(env, 4);
      if (a) (*env)->SetLongArrayRegion(env, a, 0, 4, v);
    }
  }
  if (po) (*env)->ReleaseStringUTFChars(env, combined, po);
  if (pt) (*env)->ReleaseStringUTFChars(env, transcript, pt);
  if (pg) (*env)->ReleaseStringUTFChars(env, genomic, pg);
  return a;
}

/* package utils.errorprofile;  static native int[] errorTool(long ctx, boolean multi, String bam, int maxReadLength,
 * boolean inferQualities);  -- the whole body of ErrorProfiling.inferErrorProfile (:96-621, without the plot): record loop
 * on the GPU(s) and the six output files next to the BAM.  Returns the eight run counters (numReadsProcessed, unmapped,
 * duplicates, startZero, indelRead, skippedReads, longerIndels, totalBasesChecked). */
JNIEXPORT jintArray JNICALL Java_utils_errorprofile_NativeErrorProfile_errorTool(JNIEnv* env, jclass c, jlong h, jboolean multi,
                                                                                 jstring bam, jint maxReadLength,
                                                                                 jboolean inferQualities) {
  (void)c;
  if (!bam || maxReadLength <= 0) { throw_illegal(env, "errorTool: bam must not be null, maxReadLength must be positive"); return NULL; }
  const char* pb = (*env)->GetStringUTFChars(env, bam, NULL);
  jintArray a = NULL;
  if (pb) {
    ps_profile_opts o;
    ps_fault fault;
    int32_t ctr[PS_PC_COUNT];
    memset(&o, 0, sizeof o);
    memset(&fault, 0, sizeof fault);
    o.max_read_length = (uint32_t)maxReadLength;
    o.infer_qualities = inferQualities ? 1u : 0u;
    const int st = multi ? ps_multi_error_bam((ps_multi*)(intptr_t)h, pb, &o, ctr, &fault)
                         : ps_error_bam((ps_ctx*)(intptr_t)h, pb, &o, ctr, &fault);
    if (st != PS_OK) {
      if (multi) throw_message(env, ps_multi_last_error((ps_multi*)(intptr_t)h), st);
      else throw_status(env, (ps_ctx*)(intptr_t)h, st);
    } else {
      a = (*env)->NewIntArray(env, PS_PC_COUNT);
      if (a) (*env)->SetIntArrayRegion(env, a, 0, PS_PC_COUNT, (const jint*)ctr);
    }
    (*env)->ReleaseStringUTFChars(env, bam, pb);
  }
  return a;
}
