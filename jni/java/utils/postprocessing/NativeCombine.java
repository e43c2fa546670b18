package utils.postprocessing;

/**
 * Java side of jni/parasuite_jni.c for the `comb` tool (CombineGenomeTranscript.java): the whole body of
 * CombineGenomeTranscript.combine (:36-596) + printReadsToBamFile (:598-666) in one host-side call.
 */
public final class NativeCombine {
    static {
        System.loadLibrary("parasuite_jni");
    }

    private NativeCombine() {
    }

    /** Returns mappedReads, splicedReads, missedTranscriptAlignments, liftedRecords. */
    public static native long[] combine(String genomicBam, String transcriptBam, String combinedBam);
}
