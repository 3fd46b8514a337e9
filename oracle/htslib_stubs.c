/*
 * oracle/htslib_stubs.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference's cconsenrich extension declares a handful of htslib entry points
 * (cconsenrich.pyx:26-74) that are used only by its BAM-sampling helpers
 * (cisAlignmentPairedEnd, cgetFragmentLength..., cconsenrich.pyx:4178-4650).  None of
 * them is on the filter / smoother / ECM hot path.  Building the vendored htslib just to
 * satisfy the dynamic linker would take minutes, so oracle/_ref links these aborting
 * stubs instead.  Calling any of them is a bug in the test harness.
 */
#include <stdio.h>
#include <stdlib.h>

#define CB200_STUB(name)                                                            \
    void *name(void) {                                                              \
        fprintf(stderr, "oracle/_ref: htslib stub '%s' called (not on hot path)\n", \
                #name);                                                             \
        abort();                                                                    \
        return NULL;                                                                \
    }

CB200_STUB(hts_set_threads)
CB200_STUB(hts_idx_destroy)
CB200_STUB(hts_itr_destroy)
CB200_STUB(hts_open)
CB200_STUB(hts_close)
CB200_STUB(sam_hdr_read)
CB200_STUB(sam_hdr_destroy)
CB200_STUB(bam_init1)
CB200_STUB(bam_destroy1)
CB200_STUB(sam_read1)
CB200_STUB(sam_index_load)
CB200_STUB(sam_hdr_name2tid)
CB200_STUB(sam_itr_queryi)
CB200_STUB(hts_itr_next)
CB200_STUB(bam_endpos)
CB200_STUB(bam_cigar2qlen)
CB200_STUB(hts_itr_multi_next)
CB200_STUB(hts_idx_load)
CB200_STUB(hts_log)
