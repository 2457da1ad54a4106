/* TEST INFRASTRUCTURE ONLY.
 * The reference's coal.cpp references htslib entry points for its bcf/bam front-ends
 * (coal.cpp:595-2069), which the tmp/tmp hot path never calls.  The vendored htslib
 * cannot be built in this image (no lzma/bz2/curl headers), so the stage-level probe
 * library resolves those symbols to these aborting stubs (ctypes loads with RTLD_NOW). */
#include <stdio.h>
#include <stdlib.h>
#define STUB(name) void name(void) { fprintf(stderr, "oracle/_ref: htslib symbol " #name " is stubbed (bcf/bam inputs are out of scope)\n"); abort(); }
STUB(bam_destroy1) STUB(bam_init1) STUB(bcf_destroy) STUB(bcf_get_format_values)
STUB(bcf_hdr_destroy) STUB(bcf_hdr_id2int) STUB(bcf_hdr_read) STUB(bcf_init)
STUB(bcf_is_snp) STUB(bcf_read) STUB(bcf_unpack) STUB(hts_close) STUB(hts_open)
STUB(sam_hdr_read) STUB(sam_read1)
const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";
