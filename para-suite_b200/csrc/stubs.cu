// Entry points declared in include/parasuite_b200.h whose implementation has not landed yet.
// They fail loudly (never fall back to a CPU path).
#include "internal.h"

extern "C" {
int ps_reference_load_fasta(ps_ctx* ctx, const char*) { return set_error(ctx, PS_ERR_UNSUPPORTED, "ps_reference_load_fasta: not implemented yet"); }
int ps_profile_bam(ps_ctx* ctx, const char*, const ps_profile_opts*, ps_profile_result*) { return set_error(ctx, PS_ERR_UNSUPPORTED, "ps_profile_bam: not implemented yet"); }
int ps_pileup_bam(ps_ctx* ctx, const char*, const ps_pileup_opts*, ps_pileup**) { return set_error(ctx, PS_ERR_UNSUPPORTED, "not implemented yet"); }
}
