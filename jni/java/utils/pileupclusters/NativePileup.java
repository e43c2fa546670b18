package utils.pileupclusters;

/**
 * Java side of jni/parasuite_jni.c for the `clust` tool (PileupClusters.java): the native methods the shim exports
 * under Java_utils_pileupclusters_NativePileup_*.  Contexts come from utils.errorprofile.NativeErrorProfile.create /
 * createMulti + loadReference.  INTEGRATION.md shows the two ways of patching PileupClusters.calculateReadPileups.
 */
public final class NativePileup {
    static {
        System.loadLibrary("parasuite_jni");
    }

    private NativePileup() {
    }

    /** ps_pileup_bam: the record loop PileupClusters.java:137-500 over a file; returns a handle on the cluster records. */
    public static native long pileupBam(long ctx, String bam);

    /** numReadsProcessed, skippedDueIndel, doubleStranded, nClusters, nSites, hasOpenCluster. */
    public static native long[] counters(long handle);

    /**
     * ps_pileup_next: up to maxClusters closed clusters from `first` on, 8 longs per cluster and 3 per site (layouts in
     * INTEGRATION.md); returns the number of clusters copied.
     */
    public static native int nextClusters(long handle, long first, long[] cluster64, int maxClusters, long[] site64, int maxSites);

    public static native void close(long handle);

    /**
     * The whole body of PileupClusters.calculateReadPileups (:62-584): record loop on the GPU(s) in windows, flush and the
     * six output files natively.  multi: ctx is a createMulti handle.  Returns numReadsProcessed, skippedDueIndel,
     * doubleStranded, nClusters, nSites, hasOpenCluster.
     */
    public static native long[] clustBam(long ctx, boolean multi, String bam, String out, String snpVcf, int minReadCoverage);
}
