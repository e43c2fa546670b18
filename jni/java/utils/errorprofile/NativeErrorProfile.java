package utils.errorprofile;

/**
 * Java side of jni/parasuite_jni.c for the `error` tool (ErrorProfiling.java): the native methods the shim exports
 * under Java_utils_errorprofile_NativeErrorProfile_*.  Drop this file into the reference's source tree next to
 * ErrorProfiling.java; INTEGRATION.md shows the patch of ErrorProfiling.inferErrorProfile that calls it.
 * (Not compiled in the build image of parasuite-b200: no JDK there.  The shim itself is type-checked against a stub
 * jni.h, tests/test_abi_cpu.py.)
 */
public final class NativeErrorProfile {
    static {
        System.loadLibrary("parasuite_jni");      // links libparasuite_b200.so
    }

    private NativeErrorProfile() {
    }

    /** ps_create: one context on one GPU. */
    public static native long create(int device);

    /** ps_destroy. */
    public static native void destroy(long ctx);

    /** ps_reference_load_fasta: new IndexedFastaSequenceFile(fasta) (ErrorProfiling.java:109-110); needs fasta + ".fai". */
    public static native void loadReference(long ctx, String fasta);

    /**
     * The record loop ErrorProfiling.java:146-409 over a BAM or SAM file.  Arrays are caller-allocated:
     * positionConversions maxLen*16 (index i*16 + ref*4 + read), qualityPerMismatch and its counts 16 each,
     * insertionsPerPos / deletionsPerPos maxLen each, qualityHist maxLen*256 with inferQualities (else null),
     * counters 8 (numReadsProcessed, unmapped, duplicates, startZero, indelRead, skippedReads, longerIndels,
     * totalBasesChecked).  Throws RuntimeException with the library's message (unsorted file, a record the JVM would have
     * died on, ...), IllegalArgumentException for arrays of the wrong length.
     */
    public static native void profileBam(long ctx, String bam, int maxReadLength, boolean inferQualities,
                                         int[] positionConversions, int[] qualityPerMismatch, int[] qualityPerMismatchCounts,
                                         double[] insertionsPerPos, double[] deletionsPerPos, long[] qualityHist, int[] counters);

    /** ps_create_multi: several GPUs behind one handle; null or empty: PARASUITE_B200_DEVICES ("0,1,2"), else device 0. */
    public static native long createMulti(int[] devices);

    public static native void destroyMulti(long multi);

    public static native void loadReferenceMulti(long multi, String fasta);

    /** profileBam over all devices of the handle (batches round-robin, sums added on the host). */
    public static native void profileBamMulti(long multi, String bam, int maxReadLength, boolean inferQualities,
                                              int[] positionConversions, int[] qualityPerMismatch, int[] qualityPerMismatchCounts,
                                              double[] insertionsPerPos, double[] deletionsPerPos, long[] qualityHist,
                                              int[] counters);

    /**
     * The whole body of ErrorProfiling.inferErrorProfile (:96-621 without the plot): record loop on the GPU(s) and the six
     * output files next to the BAM.  multi: ctx is a createMulti handle.  Returns the eight run counters.
     */
    public static native int[] errorTool(long ctx, boolean multi, String bam, int maxReadLength, boolean inferQualities);
}
