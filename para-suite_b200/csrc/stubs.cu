// Entry points declared in include/parasuite_b200.h whose implementation has not landed yet.
// They fail loudly (never fall back to a CPU path).
#include "internal.h"

extern "C" {
int ps_reference_load_fasta(ps_ctx* ctx, const char*) { return set_error(ctx, PS_ERR_UNSUPPORTED, "ps_reference_load_fasta: not implemented yet"); }
int ps_profile_bam(ps_ctx* ctx, const char*, const ps_profile_opts*, ps_profile_result*) { return set_error(ctx, PS_ERR_UNSUPPORTED, "ps_profile_bam: not implemented yet"); }
int ps_pileup_batch(ps_ctx* ctx, const ps_read_batch*, const ps_pileup_opts*, ps_pileup**) { return set_error(ctx, PS_ERR_UNSUPPORTED, "ps_pileup_batch: not implemented yet"); }
int ps_pileup_batch_device(ps_ctx* ctx, const ps_read_batch*, const ps_pileup_opts*, void*, ps_pileup**) { return set_error(ctx, PS_ERR_UNSUPPORTED, "not implemented yet"); }
int ps_pileup_counters_get(const ps_pileup*, ps_pileup_counters*) { return PS_ERR_UNSUPPORTED; }
int64_t ps_pileup_next(ps_pileup*, uint64_t, ps_cluster*, uint64_t, ps_site*, uint64_t) { return PS_ERR_UNSUPPORTED; }
int ps_pileup_open_cluster(ps_pileup*, ps_cluster*, ps_site*, uint64_t) { return PS_ERR_UNSUPPORTED; }
int ps_pileup_fault(const ps_pileup*, ps_fault*) { return PS_ERR_UNSUPPORTED; }
void ps_pileup_close(ps_pileup*) {}
int ps_pileup_bam(ps_ctx* ctx, const char*, const ps_pileup_opts*, ps_pileup**) { return set_error(ctx, PS_ERR_UNSUPPORTED, "not implemented yet"); }
}
